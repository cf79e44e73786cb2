"""Pins the CPU oracle (oracle/) against the golden vectors of the reference's own unit tests and, when
oracle/_ref was built from the reference sources, against the reference's compiled kernels.

Golden sources (paths under /root/reference/src):
  spline2/tests/test_multi_spline.cpp:27-41 (SymTrace), :44-80 (prefactors), :231-314 (5^3 sine grid V/G/H/L)
  QMCWaveFunctions/tests/createTestMatrix.h:36-67 + test_DiracMatrix.cpp:58-81 (3x3 inverse, logdet)
  QMCWaveFunctions/tests/test_DiracMatrix.cpp:315-366 (row update: det ratio 0.178276269185, updated inverse)
  QMCWaveFunctions/tests/test_J2_bspline.cpp:145-175 (functor u,du,d2u table), :84-88, :219-259 (logpsi, ratios)
"""
import numpy as np
import pytest

approx = pytest.approx


def test_symtrace(orc):
    assert orc.symtrace([1, 2, 3, 4.4, 1.1, 0.9], [0.1, 1.6, 1.2, 2.3, 9.4, 2.3]) == approx(29.43)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_prefactors(orc, dt):
    a, _, _ = orc.prefactors(0.1, dt)
    assert a == approx([0.1215, 0.657167, 0.221167, 0.000166667], rel=1e-5)
    a, da, d2a = orc.prefactors(0.8, dt)
    assert a == approx([0.00133333, 0.282667, 0.630667, 0.0853333], rel=1e-5)
    assert da == approx([-0.02, -0.64, 0.34, 0.32], rel=1e-5)
    assert d2a == approx([0.2, 0.4, -1.4, 0.8], rel=1e-5)


def sine_grid(N=5):
    i = np.arange(N) / N
    tpi = 2 * np.pi
    return (np.sin(tpi * i)[:, None, None] + np.sin(3 * tpi * i)[None, :, None] + np.sin(4 * tpi * i)[None, None, :])


def multi_table(coefs1, dt, nspl=1):
    from qmcpack_b200.workload import aligned_size
    npad = aligned_size(dt, nspl)
    out = np.zeros(coefs1.shape + (npad,), dt)
    for m in range(nspl):
        out[..., m] = coefs1
    return out


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_multi_bspline_periodic_5cubed(orc, dt):
    """test_multi_spline.cpp:231-314, float and double."""
    coefs = multi_table(orc.create_periodic_coefs(sine_grid()), dt)
    tol = dict(rel=2e-5, abs=2e-5) if dt == np.float32 else dict(rel=1e-8, abs=1e-8)
    v, _, _ = orc.spline_eval(coefs, 1, 0, (0, 0, 0))
    assert v[0] == approx(-3.529930688e-12, abs=1e-6)
    v, g, h = orc.spline_eval(coefs, 1, 2, (0, 0, 0))
    assert g[:, 0] == approx([6.178320809, -7.402942564, -6.178320809], **tol)
    assert h[:, 0] == approx(np.zeros(6), abs=2e-4 if dt == np.float32 else 1e-8)
    pos = (0.1, 0.2, 0.3)
    v, _, _ = orc.spline_eval(coefs, 1, 0, pos)
    assert v[0] == approx(-0.9476393279, **tol)
    v, g, h = orc.spline_eval(coefs, 1, 2, pos)
    assert v[0] == approx(-0.9476393279, **tol)
    assert g[:, 0] == approx([5.111042137, 5.989106342, 1.952244379], **tol)
    habs = 2e-4 if dt == np.float32 else 1e-7
    assert h[:, 0] == approx([-21.34557341, 1.174505743e-09, -1.1483271e-09, 133.9204891, -2.15319293e-09,
                              34.53786329], rel=2e-5, abs=habs)
    v, g, l = orc.spline_eval(coefs, 1, 1, pos)
    assert v[0] == approx(-0.9476393279, **tol)
    assert g[:, 0] == approx([5.111042137, 5.989106342, 1.952244379], **tol)
    assert l[0, 0] == approx(147.1127789, rel=2e-5)


def test_batched_positions_match_single(orc):
    """test_multi_spline.cpp:316-349: the batched position API returns the single-position numbers."""
    coefs = multi_table(orc.create_periodic_coefs(sine_grid()), np.float64)
    G = np.eye(3)
    psi, dpsi, d2psi = orc.r2r_vgl(coefs, G, 1, [[0.1, 0.2, 0.3], [0.3, 0.1, 0.2], [0.1, 0.2, 0.3]])
    assert psi[0, 0] == approx(-0.9476393279)
    assert psi[2, 0] == approx(-0.9476393279)
    assert dpsi[0, 0, 1] == approx(5.989106342)
    assert d2psi[2, 0] == approx(147.1127789)


def test_dirac_matrix_inverse(orc):
    a = np.array([[2.3, 4.5, 2.6], [0.5, 8.5, 3.3], [1.8, 4.4, 4.9]])
    b = np.array([[0.6159749342, -0.2408954682, -0.1646081192], [0.07923894288, 0.1496231042, -0.1428117337],
                  [-0.2974298429, -0.04586322768, 0.3927890292]])
    inv, logdet = orc.invert_transpose(np.ascontiguousarray(a.T))
    assert inv == approx(b, rel=1e-8)
    assert logdet.real == approx(3.78518913425)
    assert np.exp(1j * logdet.imag) == approx(1.0, abs=1e-12)  # phase is defined modulo 2*pi (LogComplexApprox)
    inv, logdet = orc.invert_transpose(np.eye(3))
    assert inv == approx(np.eye(3))
    assert abs(logdet) == approx(0.0, abs=1e-14)


INV_A = np.array([2, 5, 8, 7, 5, 2, 2, 8, 7, 5, 6, 6, 5, 4, 4, 8], float).reshape(4, 4)
INV_A_LITERAL = np.array([-0.08247423, -0.26804124, 0.26804124, 0.05154639, 0.18556701, -0.89690722, 0.39690722, 0.13402062,
                          0.24742268, -0.19587629, 0.19587629, -0.15463918, -0.29896907, 1.27835052, -0.77835052,
                          0.06185567]).reshape(4, 4)


def test_invert_transpose_4x4_literals(orc):
    """test_DiracMatrixInverterCUDA.cpp:63-110 (same matrix and values as test_cuBLAS_LU.cpp:67-110,460-520): inverse
    transpose and the complex log value, whose phase counts pi per negative pivot WITHOUT reduction (2 pi here)"""
    inv, logdet = orc.invert_transpose(INV_A, lda=4)
    assert inv[:, :4] == pytest.approx(INV_A_LITERAL, abs=2e-8)
    assert logdet.real == pytest.approx(5.267858159063328, rel=1e-12)
    assert logdet.imag == pytest.approx(6.283185307179586, rel=1e-12)


def test_complex_log_determinant_literal(orc):
    """test_cuBLAS_LU.cpp:112-166 (log value of the LU factors) for the matrix of :295-310: log|det| = 5.603777579195571 and
    a phase of -6.1586603331188225, i.e. +0.1245249740607652 modulo 2 pi (the reference does not reduce the phase)"""
    m = np.array([2.0 + 0.1j, 5.0 + 0.1j, 8.0 + 0.5j, 7.0 + 1.0j, 5.0 + 0.1j, 2.0 + 0.2j, 2.0 + 0.1j, 8.0 + 0.5j, 7.0 + 0.2j,
                  5.0 + 1.0j, 6.0 - 0.2j, 6.0 - 0.2j, 5.0 + 0.0j, 4.0 - 0.1j, 4.0 - 0.6j, 8.0 - 2.0j]).reshape(4, 4)
    inv, logdet = orc.invert_transpose(m, lda=4)
    assert logdet.real == pytest.approx(5.603777579195571, rel=1e-13)
    assert np.exp(1j * logdet.imag) == pytest.approx(np.exp(-6.1586603331188225j), abs=1e-12)
    assert inv[:, :4] == pytest.approx(np.linalg.inv(m).T, abs=1e-13)


def test_dirac_matrix_update_row(orc):
    """test_DiracMatrix.cpp:315-366 with delay rank 1."""
    a = np.array([[2.3, 4.5, 2.6], [0.5, 8.5, 3.3], [1.8, 4.4, 4.9]])
    a_inv, _ = orc.invert_transpose(np.ascontiguousarray(a.T))
    eng = orc.du(3, 1)
    v = np.array([1.9, 2.0, 3.1])
    row = eng.get_inv_row(a_inv, 0)
    ratio = float(v @ row)
    assert ratio == approx(0.178276269185)
    eng.accept_row(a_inv, 0, v, ratio)
    b = np.array([[3.455170657, -1.35124809, -0.9233316353], [0.05476311768, 0.1591951095, -0.1362710138],
                  [-2.235099338, 0.7119205298, 0.9105960265]])
    assert a_inv == approx(b, rel=1e-8)


@pytest.mark.parametrize("k", [1, 2, 4, 8])
@pytest.mark.parametrize("batched", [False, True])
def test_delayed_update_equals_fresh_inverse(orc, k, batched):
    """Procedure of test_DiracDeterminantBatched.cpp:262-470: after any accept sequence the delayed engine must equal
    a fresh LU inverse of the explicitly updated matrix; in batched mode rejected moves are pseudo-accepted."""
    rng = np.random.default_rng(5 + k)
    n = 24
    psiM = rng.normal(size=(n, n))
    ainv, logdet0 = orc.invert_transpose(psiM, lda=32)
    eng = orc.du(n, k)
    logdet = logdet0.real
    for move in range(3 * n):
        r = move % n
        new = rng.normal(size=n)
        row = eng.get_inv_row(ainv, r)
        ratio = float(row @ new)
        accept = rng.random() < 0.6
        if accept:
            eng.accept_row(ainv, r, new, ratio)
            psiM[r] = new
            logdet += np.log(abs(ratio))
        elif batched and k > 1:
            eng.pseudo_accept_row(ainv, r)
    eng.update_inv_mat(ainv)
    fresh, ld = orc.invert_transpose(psiM, lda=32)
    assert ainv[:, :n] == approx(fresh[:, :n], rel=1e-7, abs=1e-9)
    assert logdet == approx(ld.real, rel=1e-9)


J2_UD_TEST = [0.02904699284, -0.1004179, -0.1752703883, -0.2232576505, -0.2728029201, -0.3253286875, -0.3624525145,
              -0.3958223107, -0.4268582166, -0.4394531176]
J2_VALS = [(0.00, 0.1374071801, -0.5, 0.7866949593), (0.60, -0.04952403966, -0.1706645865, 0.3110897524),
           (1.20, -0.121361995, -0.09471371432, 0.055337302), (1.80, -0.1695590431, -0.06815900213, 0.0331784053),
           (2.40, -0.2058414025, -0.05505192964, 0.01049597156), (3.00, -0.2382237097, -0.05422744821, -0.002401552969),
           (3.60, -0.2712606182, -0.05600918024, -0.003537553803), (4.20, -0.3047843679, -0.05428535477, 0.0101841028),
           (4.80, -0.3347515004, -0.04506573714, 0.01469003611), (5.40, -0.3597048574, -0.03904232165, 0.005388015505),
           (6.00, -0.3823503292, -0.03657502025, 0.003511355265), (6.60, -0.4036800017, -0.03415678101, 0.007891305516),
           (7.20, -0.4219818468, -0.02556305518, 0.02075444724), (7.80, -0.4192355508, 0.06799438701, 0.3266190181),
           (8.40, -0.3019238309, 0.32586994, 0.2880861726), (9.00, -0.09726352421, 0.2851358014, -0.4238666348),
           (9.60, -0.006239062395, 0.04679296796, -0.2339648398), (10.20, 0, 0, 0), (10.80, 0, 0, 0), (11.40, 0, 0, 0)]


def test_bspline_functor_table(orc):
    """test_J2_bspline.cpp:145-175 (u-d functor, cusp -1/2, rcut 10)."""
    r = np.array([v[0] for v in J2_VALS])
    u, du, d2u = orc.functor_eval(J2_UD_TEST, 10.0, -0.5, r)
    assert u == approx([v[1] for v in J2_VALS], rel=1e-6, abs=1e-9)
    assert du == approx([v[2] for v in J2_VALS], rel=1e-6, abs=1e-9)
    assert d2u == approx([v[3] for v in J2_VALS], rel=1e-6, abs=1e-9)


J1_C_TEST = [-0.2032153051, -0.1625595974, -0.143124599, -0.1216434956, -0.09919771951, -0.07111729038, -0.04445345869,
             -0.02135082917]
J1_VALS = [(0.00, -0.1896634025, 0, 0.06586224647), (0.60, -0.1804990512, 0.02606308248, 0.02101469513),
           (1.20, -0.1637586749, 0.0255799351, -0.01568108497), (1.80, -0.1506226948, 0.01922435549, -0.005504180392),
           (2.40, -0.1394848415, 0.01869442683, 0.001517191423), (3.00, -0.128023472, 0.01946283614, 0.00104417293),
           (3.60, -0.1161729491, 0.02009651096, 0.001689229059), (4.20, -0.1036884223, 0.02172284322, 0.003731878464),
           (4.80, -0.08992443283, 0.0240346508, 0.002736384838), (5.40, -0.07519614609, 0.02475121662, -0.000347832122),
           (6.00, -0.06054074137, 0.02397053075, -0.001842295859), (6.60, -0.04654631918, 0.0225837382, -0.002780345968),
           (7.20, -0.03347994129, 0.02104406699, -0.00218107833), (7.80, -0.0211986378, 0.01996899618, -0.00173646255),
           (8.40, -0.01004416026, 0.01635533409, -0.01030907776), (9.00, -0.002594125744, 0.007782377232, -0.01556475446),
           (9.60, -0.0001660240476, 0.001245180357, -0.006225901786), (10.20, 0, 0, 0), (10.80, 0, 0, 0), (11.40, 0, 0, 0)]


def test_j1_functor_table_logpsi_and_ratios(orc):
    """test_J1_bspline.cpp: the electron-ion functor (8 coefficients, rcut 10, cusp 0) table :213-243; log psi =
    0.3160552244 for one ion at (2,0,0) and electrons at (1,0,0), (0,0,0) (:131-132, J1 = exp(-sum u)); the ratios of moving
    either electron to (0.3, 0.2, 0.5), 0.9819208747 and 1.0040884258 (:282-292).  Open boundaries: plain distances."""
    r = np.array([v[0] for v in J1_VALS])
    u, du, d2u = orc.functor_eval(J1_C_TEST, 10.0, 0.0, r)
    assert u == approx([v[1] for v in J1_VALS], rel=1e-6, abs=1e-9)
    assert du == approx([v[2] for v in J1_VALS], rel=1e-6, abs=1e-9)
    assert d2u == approx([v[3] for v in J1_VALS], rel=1e-6, abs=1e-9)
    ion = np.array([2.0, 0.0, 0.0])
    elec = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    newpos = np.array([0.3, 0.2, 0.5])
    d_old = np.linalg.norm(elec - ion, axis=1)
    d_new = np.linalg.norm(newpos - ion)
    u_old, _, _ = orc.functor_eval(J1_C_TEST, 10.0, 0.0, d_old)
    u_new, _, _ = orc.functor_eval(J1_C_TEST, 10.0, 0.0, np.array([d_new]))
    assert -u_old.sum() == approx(0.3160552244, rel=1e-6)
    assert np.exp(u_old[0] - u_new[0]) == approx(0.9819208747, rel=1e-6)
    assert np.exp(u_old[1] - u_new[0]) == approx(1.0040884258, rel=1e-6)


def test_j2_move_ratios_and_accept(orc):
    """test_J2_bspline.cpp:213-246: electrons (up, down) at (1,0,0), (0,0,0), open boundaries; moving either to
    (0.3, 0.2, 0.5) gives the ratios 0.9522052017 / 0.9871985577, a virtual move of electron 1 to (0.2, 0.5, 0.3) gives
    0.9989268241, and after accepting the move of electron 1 the log value is 0.0883791773"""
    def u(r):
        return orc.functor_eval(J2_UD_TEST, 10.0, -0.5, np.array([r]))[0][0]
    r0, r1 = np.array([1.0, 0.0, 0.0]), np.zeros(3)
    new, new2 = np.array([0.3, 0.2, 0.5]), np.array([0.2, 0.5, 0.3])
    d01 = np.linalg.norm(r0 - r1)
    assert np.exp(u(d01) - u(np.linalg.norm(new - r1))) == approx(0.9522052017, rel=1e-6)
    assert np.exp(u(d01) - u(np.linalg.norm(new - r0))) == approx(0.9871985577, rel=1e-6)
    assert np.exp(u(d01) - u(np.linalg.norm(new2 - r0))) == approx(0.9989268241, rel=1e-6)
    assert -u(np.linalg.norm(new - r0)) == approx(0.0883791773, rel=1e-6)


def _two_electron_system():
    # open-boundary test geometry emulated by a huge cubic cell (rcut = 10 << L/2)
    L = 400.0
    table = np.zeros((4, 4, 4, 8), np.float64)  # 1^3 grid, constant orbitals (never used in this test)
    return dict(n_up=1, n_dn=1, lattice=np.eye(3) * L, coefs=[table, table],
                j2=dict(uu=J2_UD_TEST, ud=J2_UD_TEST, rcut=10.0))


def test_j2_logpsi_two_electrons(orc):
    """test_J2_bspline.cpp:84-88: electrons at (1,0,0),(0,0,0): log psi = 0.1012632641, KE = -0.1616624771."""
    import oracle_lib
    sysd = _two_electron_system()
    # constant non-zero orbital tables so the 1x1 determinants are regular; J2 is what is checked
    sysd["coefs"] = [np.ones((4, 4, 4, 8)), np.ones((4, 4, 4, 8))]
    vmc = oracle_lib.OracleVMC(orc, sysd, nw=1, delay_rank=1)
    vmc.set_positions(np.array([[[1.0, 0, 0], [0, 0, 0]]]))
    vmc.recompute()
    Uat, dUat, d2Uat = vmc.j2_state(0)
    logpsi_j2 = -0.5 * Uat.sum()
    assert logpsi_j2 == approx(0.1012632641)
    ke = -0.5 * ((dUat ** 2).sum() + d2Uat.sum())
    assert ke == approx(-0.1616624771)


def test_mt19937_stream(orc):
    """std::mt19937 KAT (10000th output of a default-seeded engine is 4123659995, [rand.predef]) and the
    uniform_real_distribution_as_boost mapping (Utilities/StdRandom.h:43-47)."""
    raw = orc.rng_raw(5489, 10000)
    assert int(raw[-1]) == 4123659995
    u = orc.rng_uniform(5489, 4)
    assert u == approx(raw[:4].astype(np.float64) / 4294967296.0, rel=0, abs=0)


def test_box_muller(orc):
    """RandomSeqGenerator.h:33-52 restated in numpy from the same raw stream."""
    n = 7
    u = orc.rng_uniform(911, 8)
    g = orc.rng_gauss(911, n)
    exp = []
    for i in range(0, 8, 2):
        t1 = np.sqrt(-2.0 * np.log(1.0 - (1.0 - np.finfo(np.float64).eps) * u[i]))
        t2 = 2.0 * np.pi * u[i + 1]
        exp += [t1 * np.cos(t2), t1 * np.sin(t2)]
    assert g == approx(exp[:n], rel=1e-14)


# ---------------------------------------------------------------------------------------------------
# restatement vs the reference's own compiled kernels (oracle/_ref)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_port_matches_reference_spline_kernels(orc, orc_ref, dt):
    from qmcpack_b200.workload import random_table
    coefs = random_table((7, 5, 6), 19, dt, seed=3)
    rng = np.random.default_rng(0)
    for pos in rng.random((20, 3)):
        for which in (0, 1, 2):
            a = orc.spline_eval(coefs, 19, which, pos)
            b = orc_ref.spline_eval(coefs, 19, which, pos)
            for x, y in zip(a, b):
                # same association order; only FMA contraction choices of the two compilations may differ
                assert x[..., :19] == approx(y[..., :19], rel=5e-6 if dt == np.float32 else 1e-13,
                                             abs=1e-4 if dt == np.float32 else 1e-11)


def test_port_matches_reference_einspline_solver(orc, orc_ref):
    data = np.random.default_rng(1).normal(size=(6, 5, 7))
    assert orc.create_periodic_coefs(data) == approx(orc_ref.create_periodic_coefs(data), rel=1e-12, abs=1e-13)


def test_port_matches_reference_inverse(orc, orc_ref):
    a = np.random.default_rng(2).normal(size=(40, 40))
    i1, l1 = orc.invert_transpose(a)
    i2, l2 = orc_ref.invert_transpose(a)
    assert i1 == approx(i2, rel=1e-9, abs=1e-11)
    assert l1.real == approx(l2.real, rel=1e-12)
    assert np.cos(l1.imag) == approx(np.cos(l2.imag), abs=1e-12)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("k", [1, 4, 8])
def test_port_matches_reference_delayed_update(orc, orc_ref, dt, k):
    """Same accept sequence through the restated engine and through qmcplusplus::DelayedUpdate<T>."""
    rng = np.random.default_rng(11)
    n = 16
    psiM = (2 * np.eye(n) + 0.3 * rng.normal(size=(n, n))).astype(dt)  # well conditioned: only rounding differs
    a1, _ = orc.invert_transpose(psiM)
    a2 = a1.copy()
    e1, e2 = orc.du(n, k, dt), orc_ref.du(n, k, dt)
    tol = dict(rel=2e-4, abs=2e-5) if dt == np.float32 else dict(rel=1e-8, abs=1e-9)
    for move in range(2 * n + 3):
        r = move % n
        new = (2 * np.eye(n)[r] + 0.3 * rng.normal(size=n)).astype(dt)
        r1, r2 = e1.get_inv_row(a1, r), e2.get_inv_row(a2, r)
        assert r1 == approx(r2, **tol)
        ratio = float(r2 @ new)
        if rng.random() < 0.7 and abs(ratio) > 0.2:  # keep the walk well conditioned: rounding is what differs
            e1.accept_row(a1, r, new, ratio)
            e2.accept_row(a2, r, new, ratio)
    e1.update_inv_mat(a1)
    e2.update_inv_mat(a2)
    scale = float(np.abs(a2).max())
    assert a1 == approx(a2, rel=tol["rel"], abs=tol["abs"] * max(1.0, scale))


# ---------------------------------------------------------------------------------------------------------------------
# complex value types: restatement vs the reference-compiled kernels (DelayedUpdate<std::complex<T>>, DiracMatrix, zgetrf)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
@pytest.mark.parametrize("k", [1, 4, 8])
def test_complex_delayed_update_port_matches_reference(orc, orc_ref, dt, k):
    if orc_ref is None:
        pytest.skip("oracle/_ref not built")
    n = 24
    rng = np.random.default_rng(5 + k)
    a = (2 * np.eye(n) + 0.1 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))).astype(dt)
    invs, engs, lds = [], [], []
    for o in (orc, orc_ref):
        inv, ld = o.invert_transpose(a)
        invs.append(inv)
        lds.append(ld)
        engs.append(o.du(n, k, dt))
    tol = 1e-12 if dt == np.complex128 else 1e-4
    assert np.abs(invs[0] - invs[1]).max() < tol
    assert abs(lds[0] - lds[1]) < 1e-10
    # (A^-1)^T A^T = 1 for the complex inverse
    assert np.abs(invs[0].T.astype(np.complex128) @ a.astype(np.complex128) - np.eye(n)).max() < (1e-12 if dt == np.complex128 else 1e-4)
    for move in range(2 * k + 3):
        row = (3 * move) % n
        v = (2 * np.eye(n)[row] + 0.1 * (rng.normal(size=n) + 1j * rng.normal(size=n))).astype(dt)
        rows = [e.get_inv_row(inv, row) for e, inv in zip(engs, invs)]
        assert np.abs(rows[0] - rows[1]).max() < tol * 10
        ratio = (rows[0].astype(np.complex128) @ v.astype(np.complex128)).item()
        for e, inv in zip(engs, invs):
            e.accept_row(inv, row, v, ratio)
        a[row] = v
    for e, inv in zip(engs, invs):
        e.update_inv_mat(inv)
    assert np.abs(invs[0] - invs[1]).max() < tol * 100
    fresh, _ = orc.invert_transpose(a)
    assert np.abs(invs[0] - fresh).max() < (1e-9 if dt == np.complex128 else 5e-3)


def test_complex_vmc_port_matches_reference(orc, orc_ref):
    """the oracle's complex-orbital VMC sweep on the reference's compiled DelayedUpdate<complex<double>>, DiracMatrix and
    spline VGH kernels: identical acceptance sequence, G/L to rounding"""
    if orc_ref is None:
        pytest.skip("oracle/_ref not built")
    import oracle_lib
    from qmcpack_b200.workload import make_system, initial_positions
    s = make_system(N=24, M=8, dtype=np.float64, L=6.0, complex_orbitals=True)
    R = initial_positions(s, 4)
    out = []
    for o in (orc, orc_ref):
        v = oracle_lib.OracleVMC(o, s, nw=4, ncrowds=2, seeds=[3, 4], tau=0.1, delay_rank=4, batched_engine=False)
        v.set_positions(R)
        v.recompute()
        log = v.sweep(3, log_accept=True)
        out.append((log,) + v.evaluate_gl())
    assert 0.2 < out[0][0].mean() < 0.98
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1:], out[1][1:]):
        assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(b).max())
    assert np.abs(out[0][3].imag).max() > 1e-3
