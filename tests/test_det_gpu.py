"""GPU parity: delayed-update determinant engine through the C ABI vs the CPU oracle, following the procedure of the
reference's test_DiracDeterminantBatched.cpp:262-470 (fixed matrices instead of a spline SPOSet: "FakeSPO")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


def tiny_system(n, dt):
    """a crowd needs an SPOSet handle per determinant; the determinant tests feed orbital rows directly"""
    from qmcpack_b200.workload import random_table
    dt = np.dtype(dt)
    if dt.kind == "c":  # complex determinants sit on SplineC2C tables (2n real components + twist vectors)
        t = random_table((4, 4, 4), 2 * n, np.float32 if dt == np.complex64 else np.float64, seed=1)
        k = np.tile([0.1, 0.2, 0.3], (n, 1))
        return dict(n_up=n, n_dn=n, lattice=np.eye(3) * 4.0, coefs=[t, t], kpts=[k, k])
    t = random_table((4, 4, 4), n, dt, seed=1)
    return dict(n_up=n, n_dn=n, lattice=np.eye(3) * 4.0, coefs=[t, t])


def normal(rng, shape, dt):
    """standard normal samples of a real or complex dtype"""
    dt = np.dtype(dt)
    if dt.kind == "c":
        return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(dt)
    return rng.normal(size=shape).astype(dt)


def test_reference_3x3_literals(api):
    """test_DiracMatrix.cpp:58-81 and :315-366 literals through the CUDA engine (delay rank 1)"""
    a = np.array([[2.3, 4.5, 2.6], [0.5, 8.5, 3.3], [1.8, 4.4, 4.9]])
    crowd = api.Crowd(tiny_system(3, np.float64), nw=2, delay_rank=1)
    psiM = np.stack([a.T, a.T])  # the test inverts a_T: a_inv = (a_T^-1)^T
    crowd.det_recompute_from_matrices(0, psiM)
    inv, logdet = crowd.det_mw_completeUpdates(0)
    b = np.array([[0.6159749342, -0.2408954682, -0.1646081192], [0.07923894288, 0.1496231042, -0.1428117337],
                  [-0.2974298429, -0.04586322768, 0.3927890292]])
    assert inv[0] == pytest.approx(b, rel=1e-8)
    assert logdet[0, 0] == pytest.approx(3.78518913425)
    v = np.array([1.9, 2.0, 3.1])
    phi = np.zeros((5, 2, 3))
    phi[0, :, :] = v
    crowd.det_set_phi_vgl(0, phi)
    ratios, _ = crowd.det_mw_ratioGrad(0, 0, from_phi=True)
    assert ratios[0] == pytest.approx(0.178276269185)
    crowd.det_mw_accept_rejectRow(0, 0, [1, 0])
    inv, logdet = crowd.det_mw_completeUpdates(0)
    b2 = np.array([[3.455170657, -1.35124809, -0.9233316353], [0.05476311768, 0.1591951095, -0.1362710138],
                   [-2.235099338, 0.7119205298, 0.9105960265]])
    assert inv[0] == pytest.approx(b2, rel=1e-8)
    assert inv[1] == pytest.approx(b, rel=1e-8)  # rejected walker untouched


def test_invert_transpose_4x4_literals(api):
    """mw_invertTranspose literals of test_DiracMatrixInverterCUDA.cpp:63-132 (batch sizes 1, 2, 3 give the same result):
    inverse transpose and log value (5.267858159063328, 2 pi) through mw_recompute's FP64 LU"""
    a = np.array([2, 5, 8, 7, 5, 2, 2, 8, 7, 5, 6, 6, 5, 4, 4, 8], float).reshape(4, 4)
    want = np.array([-0.08247423, -0.26804124, 0.26804124, 0.05154639, 0.18556701, -0.89690722, 0.39690722, 0.13402062,
                     0.24742268, -0.19587629, 0.19587629, -0.15463918, -0.29896907, 1.27835052, -0.77835052,
                     0.06185567]).reshape(4, 4)
    for nw in (1, 2, 3):
        crowd = api.Crowd(tiny_system(4, np.float64), nw=nw, delay_rank=2)
        crowd.det_recompute_from_matrices(0, np.stack([a] * nw))
        inv, logdet = crowd.det_mw_completeUpdates(0)
        for iw in range(nw):
            assert inv[iw][:, :4] == pytest.approx(want, abs=2e-8)
            assert logdet[iw, 0] == pytest.approx(5.267858159063328, rel=1e-12)
            # the reference accumulates pi per negative pivot without reducing the phase; any representative of the
            # same phase modulo 2 pi is the same determinant sign
            assert np.cos(logdet[iw, 1]) == pytest.approx(1.0, abs=1e-12)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("n", [1, 5, 31, 33, 70, 130, 384])
def test_batched_inverse_and_logdet_match_lapack(api, dt, n):
    """mw_invertTranspose (DiracMatrixInverterCUDA.hpp:306-369) on the library's own blocked Gauss-Jordan kernels
    (csrc/inverse.cuh): inverse transpose, log|det| and the determinant's phase against LAPACK (numpy) in double, for
    panel-aligned and ragged sizes, real and complex; rows are permuted per walker so that pivoting is exercised (a
    different pivot order in every walker) and one walker carries a negative determinant"""
    rng = np.random.default_rng(100 + n)
    nw = 5 if n < 384 else 3
    psiM = normal(rng, (nw, n, n), dt)
    psiM += 0.5 * np.sqrt(n) * np.eye(n, dtype=dt)  # conditioned so that float results are meaningful too
    for iw in range(nw):
        psiM[iw] = psiM[iw][rng.permutation(n)]
    crowd = api.Crowd(tiny_system(n, dt), nw=nw, delay_rank=1)
    crowd.det_recompute_from_matrices(0, psiM)
    inv, logdet = crowd.det_mw_completeUpdates(0)
    ref = np.linalg.inv(psiM.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64)).transpose(0, 2, 1)
    sign, logabs = np.linalg.slogdet(psiM.astype(ref.dtype))
    single = np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64))
    tol = 2e-5 if single else 1e-10  # the inversion itself is FP64; single = the cast of the result
    for iw in range(nw):
        assert np.abs(inv[iw][:, :n] - ref[iw]).max() <= tol * np.abs(ref[iw]).max()
        assert logdet[iw, 0] == pytest.approx(logabs[iw], rel=1e-11, abs=1e-11)
        assert np.exp(1j * logdet[iw, 1]) == pytest.approx(sign[iw], abs=1e-9)
    assert (np.real(sign) < 0).any() or np.dtype(dt).kind == "c" or n == 1


def test_singular_matrix_is_reported(api):
    """getrf's info != 0 -> exception (DiracMatrixInverterCUDA.hpp:180-190 checks the infos and throws)"""
    a = np.ones((2, 6, 6))
    a[0] += np.eye(6)
    crowd = api.Crowd(tiny_system(6, np.float64), nw=2, delay_rank=1)
    with pytest.raises(RuntimeError, match="singular"):
        crowd.det_recompute_from_matrices(0, a)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_own_inverse_and_cublas_route_agree(api, dt):
    """the timing hook runs both routes on the same Slater matrices and leaves the crowd recomputed"""
    from qmcpack_b200 import workload
    s = workload.make_system(N=48, M=8, dtype=np.float64, L=6.0, complex_orbitals=np.dtype(dt).kind == "c")
    crowd = api.Crowd(s, nw=4, delay_rank=4)
    crowd.set_positions(workload.initial_positions(s, 4))
    crowd.mw_recompute()
    a0, l0 = crowd.det_mw_completeUpdates(0)
    for method in (1, 2):
        assert crowd.det_time_inverse(0, method, reps=1) > 0
        a1, l1 = crowd.det_mw_completeUpdates(0)
        assert np.abs(a1 - a0).max() <= 1e-9 * np.abs(a0).max()
        assert np.allclose(l1[:, 0], l0[:, 0], rtol=1e-12, atol=1e-12)
        assert np.allclose(np.exp(1j * l1[:, 1]), np.exp(1j * l0[:, 1]), atol=1e-9)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("n,k", [(24, 1), (24, 2), (24, 8), (70, 16), (192, 32), (130, 64)])
def test_delayed_update_sequence_matches_oracle(api, orc, dt, n, k):
    """random accept/reject sequence, every walker with its own flags: inverse rows, ratios, gradients at every move and
    the flushed inverse / log-determinant at the end must match the oracle's batched engine (pseudo-accepts included)
    and a fresh inverse of the explicitly updated matrix."""
    run_delayed_update_sequence(api, orc, dt, n, k, nmoves=2 * n + 3)


# the shapes bench.py runs: NiO-a64 (n = 384, k = 32: three 128-row tiles x three column pieces of the tcgen05 flush, six
# 64-row tiles of the DMMA flush), NiO-a256 (n = 1536, k = 32) and NiO-a128 (n = 768 complex double, k = 64).  Four full
# flushes plus a partial one, every move checked.
@pytest.mark.parametrize("dt,n,k", [(np.float32, 384, 32), (np.float64, 384, 32), (np.complex64, 384, 32),
                                    (np.complex128, 384, 32), (np.float32, 1536, 32), (np.float64, 1536, 32),
                                    (np.complex128, 768, 64), (np.float32, 192, 32), (np.float32, 384, 64)],
                         ids=["a64-f32", "a64-f64", "a64-c64", "a64-c128", "a256-f32", "a256-f64", "a128-c128", "a32-f32",
                              "a64-f32-k64"])
def test_delayed_update_benchmarked_shapes(api, orc, dt, n, k):
    run_delayed_update_sequence(api, orc, dt, n, k, nmoves=4 * k + 5, nw=3)


def run_delayed_update_sequence(api, orc, dt, n, k, nmoves, nw=5):
    rng = np.random.default_rng(100 + n + k)
    crowd = api.Crowd(tiny_system(n, dt), nw=nw, delay_rank=k)
    amp = 0.5 / np.sqrt(n)  # keeps the matrices well conditioned so that only rounding separates GPU and CPU
    psiM = (2 * np.eye(n) + amp * normal(rng, (nw, n, n), dt)).astype(dt)
    dpsiM = normal(rng, (nw, n, n, 3), dt)
    d2psiM = normal(rng, (nw, n, n), dt)
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    crowd.det_recompute_from_matrices(1, psiM, dpsiM, d2psiM)
    lda = orc.aligned_size(dt, n)
    ainv, logdet, eng = [], [], []
    for iw in range(nw):
        a, ld = orc.invert_transpose(psiM[iw], lda=lda)
        ainv.append(a)
        logdet.append(ld.real)
        eng.append(orc.du(n, k, dt))
    inv0, ld0 = crowd.det_mw_completeUpdates(1)
    tol = 1e-8 if np.dtype(dt).itemsize // (2 if np.dtype(dt).kind == 'c' else 1) == 8 else 2e-4
    fp64 = tol == 1e-8
    for iw in range(nw):
        assert np.abs(inv0[iw] - ainv[iw][:, :n]).max() < tol * np.abs(ainv[iw]).max()
        assert ld0[iw, 0] == pytest.approx(logdet[iw], rel=1e-10)
    cur_dpsiM = dpsiM.copy()
    for move in range(nmoves):
        row = move % n
        grads_now = crowd.det_mw_evalGrad(1, row)
        rows_gpu = crowd.det_mw_getInvRow(1, row)
        phi = np.zeros((5, nw, n), dt)
        phi[0] = (2 * np.eye(n)[row] + amp * normal(rng, (nw, n), dt)).astype(dt)
        phi[1:] = normal(rng, (4, nw, n), dt)
        crowd.det_set_phi_vgl(1, phi)
        ratios, grads = crowd.det_mw_ratioGrad(1, row, from_phi=True)
        acc = (rng.random(nw) < 0.6).astype(np.uint8)
        for iw in range(nw):
            r = eng[iw].get_inv_row(ainv[iw], row)
            assert np.abs(rows_gpu[iw] - r).max() < tol * max(1.0, np.abs(r).max())
            g_ref = r.astype(wide) @ cur_dpsiM[iw, row].astype(wide)  # plain (unconjugated) dots, SPOSet.cpp:166-171
            assert np.abs(grads_now[iw] - g_ref).max() < tol * 50 * max(1.0, np.abs(g_ref).max())
            ratio_ref = (r.astype(wide) @ phi[0, iw].astype(wide)).item()
            assert abs(ratios[iw] - ratio_ref) < tol * 10 * max(1.0, abs(ratio_ref))
            gn_ref = (r.astype(wide) @ phi[1:4, iw].astype(wide).T) / ratio_ref
            assert np.abs(grads[iw] - gn_ref).max() < tol * 100 * max(1.0, np.abs(gn_ref).max())
            if acc[iw]:
                eng[iw].accept_row(ainv[iw], row, phi[0, iw], ratios[iw].item())
                psiM[iw, row] = phi[0, iw]
                cur_dpsiM[iw, row] = phi[1:4, iw].T
                logdet[iw] += np.log(abs(ratios[iw].item()))
            elif k > 1:
                eng[iw].pseudo_accept_row(ainv[iw], row)
        crowd.det_mw_accept_rejectRow(1, row, acc)
        assert crowd.det_delay_count(1) == (move + 1) % k
    inv, ld = crowd.det_mw_completeUpdates(1)
    for iw in range(nw):
        eng[iw].update_inv_mat(ainv[iw])
        scale = np.abs(ainv[iw]).max()
        assert np.abs(inv[iw] - ainv[iw][:, :n]).max() < tol * scale
        fresh, fld = orc.invert_transpose(psiM[iw], lda=lda)
        assert np.abs(inv[iw] - fresh[:, :n]).max() < (1e-8 if fp64 else 5e-3) * scale
        assert ld[iw, 0] == pytest.approx(fld.real, rel=1e-9 if fp64 else 1e-4, abs=1e-5)
        # phase of the determinant (imaginary part of the complex log) modulo 2 pi
        dphase = (ld[iw, 1] - fld.imag + np.pi) % (2 * np.pi) - np.pi
        assert abs(dphase) < (1e-8 if fp64 else 1e-3)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_determinant_from_spline_recompute(api, orc, dt):
    """mw_recompute through the spline SPOSet: psiM rows are orbital values at the electron positions; inverse and
    log-determinant against the oracle's LU."""
    from qmcpack_b200.workload import make_system, initial_positions
    s = make_system(N=32, M=6, dtype=dt, L=5.0, with_j1=False, with_j2=False)
    nw = 3
    crowd = api.Crowd(s, nw=nw, delay_rank=4)
    R = initial_positions(s, nw)
    crowd.set_positions(R)
    crowd.mw_recompute()
    G = np.linalg.inv(s["lattice"])
    for spin in (0, 1):
        n = 16
        inv, ld = crowd.det_mw_completeUpdates(spin)
        for iw in range(nw):
            pos = R[iw, spin * 16:(spin + 1) * 16]
            psiM, _, _ = orc.r2r_vgl(s["coefs"][spin], G, n, pos)
            ref, rld = orc.invert_transpose(np.ascontiguousarray(psiM))
            tol = 1e-8 if dt == np.float64 else 5e-3
            assert np.abs(inv[iw] - ref).max() < tol * np.abs(ref).max()
            assert ld[iw, 0] == pytest.approx(rld.real, rel=1e-9 if dt == np.float64 else 1e-4, abs=1e-4)
