"""GPU parity of the whole per-electron-move path against the oracle's restatement of VMCBatched::advanceWalkers:
identical random streams, identical acceptance sequences (FP64), statistically equal energies (mixed precision)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


LAT_GENERAL = np.array([[6.0, 0.4, 0.0], [0.3, 6.5, -0.2], [0.1, -0.3, 7.0]])
# the same lattice described by a NON-reduced basis (a1 -> a1 + 2 a0, a2 -> a2 + a0 - a1): the minimum image must come from
# the reduced basis (find_reduced_basis, LatticeAnalyzer.h:213-272; ParticleBConds3DSoa.h:339-386)
LAT_SHEARED = np.array([[1, 0, 0], [2, 1, 0], [1, -1, 1]], float) @ LAT_GENERAL


def small_system(dt, lattice=None, N=24, M=8):
    from qmcpack_b200.workload import make_system
    return make_system(N=N, M=M, dtype=dt, L=6.0, lattice=lattice)


@pytest.mark.parametrize("lattice", [None, LAT_GENERAL, LAT_SHEARED], ids=["ortho", "general", "non_reduced"])
def test_distance_rows_and_j2_ratio(api, orc, lattice):
    """SoaDistanceTableAA temp/old rows and TwoBodyJastrow::mw_ratioGrad against the oracle (orthorhombic and general cell)"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    dt = np.float64
    s = small_system(dt, lattice)
    nw, N = 4, 24
    crowd = api.Crowd(s, nw=nw, delay_rank=2)
    R = initial_positions(s, nw)
    crowd.set_positions(R)
    crowd.mw_recompute()
    rng = np.random.default_rng(3)
    iat = 13
    displ = rng.normal(size=(nw, 3)) * 0.7
    crowd.mw_makeMove(iat, displ)
    rows = crowd.dtaa_temp_rows()
    npad = orc.aligned_size(dt, N)
    for iw in range(nw):
        rsoa = np.zeros((3, npad))
        rsoa[:, :N] = R[iw].T
        new = orc.dist_row(s["lattice"], R[iw, iat] + displ[iw], rsoa, N, iat)
        old = orc.dist_row(s["lattice"], R[iw, iat], rsoa, N, iat)
        assert rows[0, iw] == pytest.approx(new[:, :N], rel=1e-12, abs=1e-12)
        mask = np.arange(N) != iat
        assert rows[1, iw][:, mask] == pytest.approx(old[:, :N][:, mask], rel=1e-12, abs=1e-12)
    # J2 state after recompute and the ratio of the proposed move against the oracle VMC object
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, delay_rank=2)
    ov.set_positions(R)
    ov.recompute()
    for iw in range(nw):
        for a, b in zip(crowd.j2_state(iw), ov.j2_state(iw)):
            assert a == pytest.approx(b, rel=1e-10, abs=1e-12)


def host_sweeps(api, orc, s, nw, k, nsteps, tau, seed=1000, use_drift=True):
    """product (host-driven C-ABI calls) and oracle on one mt19937 stream; returns both acceptance logs and objects"""
    from qmcpack_b200.workload import initial_positions
    from qmcpack_b200 import vmc_host
    import oracle_lib
    N = s["n_up"] + s["n_dn"]
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, use_drift=use_drift, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    rng = orc.rng(seed)
    log = np.zeros((nsteps, N, nw), np.uint8)
    for step in range(nsteps):
        vmc_host.advance_walkers(crowd, rng, tau=tau, use_drift=use_drift, log_accept=log[step])
    olog = ov.sweep(nsteps, log_accept=True)
    return crowd, ov, log, olog


@pytest.mark.parametrize("host_kernel", ["resident", "launch"])
@pytest.mark.parametrize("lattice", [None, LAT_GENERAL, LAT_SHEARED], ids=["ortho", "general", "non_reduced"])
@pytest.mark.parametrize("k", [1, 4])
def test_host_driven_sweep_identical_acceptance_fp64(api, orc, lattice, k, host_kernel, monkeypatch):
    """FP64: acceptance sequences identical; positions, log psi, kinetic energy, G and L agree to rounding.  The
    per-electron calls are served either by the resident walker-segment kernel (host mailboxes) or by launches per call."""
    monkeypatch.setenv("QMCB_HOST_KERNEL", host_kernel)
    s = small_system(np.float64, lattice)
    crowd, ov, log, olog = host_sweeps(api, orc, s, nw=6, k=k, nsteps=3, tau=0.1)
    assert crowd.host_kernel == (2 if host_kernel == "resident" else 1)
    assert 0.2 < olog.mean() < 0.98
    assert np.array_equal(log, olog)
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)
    lp, ke, G, L = crowd.mw_evaluateGL()
    olp, oke, oG, oL = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-8, abs=1e-8)
    assert ke == pytest.approx(oke, rel=1e-6, abs=1e-6)
    assert G == pytest.approx(oG, rel=1e-6, abs=1e-6)
    # delayed update == from scratch: recompute and compare log psi (checkGL_after_moves of the reference)
    crowd.mw_recompute()
    lp2, ke2, _, _ = crowd.mw_evaluateGL()
    assert lp2 == pytest.approx(lp, rel=1e-9, abs=1e-9)
    assert ke2 == pytest.approx(ke, rel=1e-7, abs=1e-7)


# the two implementations of the device-resident sweep (qmcb_vmc_params.sweep_kernel): 1 = boundary kernel + spline gather
# per move, 2 = the persistent walker-segment kernel (csrc/segment.cuh).  Same random stream, same decisions.
SWEEP_KERNELS = pytest.mark.parametrize("sweep_kernel", [1, 2], ids=["two_kernel", "segment_kernel"])


@SWEEP_KERNELS
@pytest.mark.parametrize("lattice", [None, LAT_GENERAL], ids=["ortho", "general"])
def test_device_driver_identical_acceptance_fp64(api, orc, sweep_kernel, lattice):
    """the device-resident sweep (on-device mt19937, Box-Muller, Metropolis test) reproduces the oracle's acceptance
    sequence in FP64 with drift, J1 and J2, delay rank 4, with and without CUDA graph replay"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = small_system(np.float64, lattice)
    nw, k, nsteps, tau, seed = 7, 4, 3, 0.1, 4242
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    for graph in (False, True):
        crowd = api.Crowd(s, nw=nw, delay_rank=k)
        crowd.set_positions(R)
        crowd.mw_recompute()
        crowd.vmc_init(tau=tau, use_drift=True, seed=seed, use_cuda_graph=graph, sweep_kernel=sweep_kernel)
        assert crowd.sweep_kernel == sweep_kernel
        log = crowd.vmc_sweep(nsteps, log_accept=True)
        assert np.array_equal(log, olog), f"graph={graph}: {np.argwhere(log != olog)[:5]}"
        assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-8, abs=1e-8)
        lp, ke, G, L = crowd.mw_evaluateGL()
        olp, oke, oG, oL = ov.evaluate_gl()
        assert lp == pytest.approx(olp, rel=1e-8, abs=1e-8)
        assert ke == pytest.approx(oke, rel=1e-6, abs=1e-6)
        assert G == pytest.approx(oG, rel=1e-6, abs=1e-6)
        na, nr = crowd.vmc_counts()
        assert (na + nr == nsteps * 24).all() and na.sum() == olog.sum()
        # delayed-update state == from-scratch recompute (checkGL_after_moves of the reference)
        crowd.mw_recompute()
        lp2, ke2, _, _ = crowd.mw_evaluateGL()
        assert lp2 == pytest.approx(lp, rel=1e-9, abs=1e-9)
        assert ke2 == pytest.approx(ke, rel=1e-7, abs=1e-7)


@pytest.mark.parametrize("N,k,nw", [(24, 1, 5), (24, 12, 5), (26, 5, 33), (40, 32, 70), (400, 32, 3), (768, 32, 2)],
                         ids=["k1", "k_eq_n", "odd_k_33w", "k_gt_n_70w", "one_box_wide", "two_boxes_a64_width"])
def test_segment_kernel_shapes_fp64(api, orc, N, k, nw):
    """the persistent walker-segment kernel over the shapes that exercise its corners: delay rank 1 (every segment is one
    move), delay rank = determinant size, segments that do not divide the determinant, more walkers than one look-back
    warp covers, a determinant wider than one TMA box (200 > 192 orbitals: two components per consumer thread) and the
    NiO-a64 width (384).  FP64: acceptance identical to the oracle, delayed-update state == recompute."""
    from qmcpack_b200.workload import make_system, initial_positions
    import oracle_lib
    kk = min(k, N // 2)
    s = make_system(N=N, M=8 if N <= 40 else 12, dtype=np.float64, L=6.0 if N <= 40 else None)
    nsteps, tau, seed = 2, 0.1, 99 + N
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=kk)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    crowd = api.Crowd(s, nw=nw, delay_rank=kk)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=tau, use_drift=True, seed=seed, use_cuda_graph=True, sweep_kernel=2)
    log = crowd.vmc_sweep(nsteps, log_accept=True)
    assert np.array_equal(log, olog), np.argwhere(log != olog)[:5]
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-8, abs=1e-8)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-8, abs=1e-7)
    crowd.mw_recompute()
    lp2, ke2, _, _ = crowd.mw_evaluateGL()
    assert lp2 == pytest.approx(lp, rel=1e-9, abs=1e-7)


def test_segment_kernel_refused_for_complex_orbitals(api):
    """complex determinants run the two-kernel path; asking for the segment kernel explicitly is an error, automatic
    selection falls back silently"""
    from qmcpack_b200.workload import make_system, initial_positions
    s = make_system(N=24, M=8, dtype=np.float64, L=6.0, complex_orbitals=True)
    crowd = api.Crowd(s, nw=3, delay_rank=4)
    crowd.set_positions(initial_positions(s, 3))
    crowd.mw_recompute()
    with pytest.raises(RuntimeError, match="complex orbitals"):
        crowd.vmc_init(tau=0.1, seed=1, sweep_kernel=2)
    crowd.vmc_init(tau=0.1, seed=1, sweep_kernel=0)
    assert crowd.sweep_kernel == 1


@pytest.mark.parametrize("ncrowds", [1, 3])
def test_compiled_host_driver_matches_oracle_multi_crowd(api, orc, ncrowds):
    """the C++ host driver above the C ABI (one thread and one mt19937 stream per crowd) against the oracle with the same
    crowd partition and seeds: identical acceptance sequences in FP64"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = small_system(np.float64)
    nw, k, nsteps, tau = 7, 4, 2, 0.1
    seeds = [11 + 5 * c for c in range(ncrowds)]
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=ncrowds, seeds=seeds, tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    # same contiguous fair division of walkers over crowds as the oracle / MCPopulation
    base, extra = divmod(nw, ncrowds)
    sizes = [base + (1 if c < extra else 0) for c in range(ncrowds)]
    crowds, off = [], 0
    spo = None
    for c in range(ncrowds):
        cr = api.Crowd(s, nw=sizes[c], delay_rank=k, spo=spo)
        spo = cr.spo  # crowds share the read-only tables like clones share the SPOSet (SplineR2R.h:82)
        cr.set_positions(R[off:off + sizes[c]])
        cr.mw_recompute()
        crowds.append(cr)
        off += sizes[c]
    drv = api.HostVMC(crowds, seeds, tau=tau, use_drift=True)
    log = drv.run(nsteps, log_accept=True)
    assert np.array_equal(log, olog)
    acc, rej = drv.counts()
    assert acc == olog.sum() and acc + rej == olog.size
    got = np.concatenate([c.positions() for c in crowds])
    assert got == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)
    up, down = drv.bytes_per_sweep()
    assert up == 24 * (nw * 3 * 8 + nw) and down == 24 * (nw * 3 * 8 * 2 + nw * 8)


@SWEEP_KERNELS
def test_device_driver_no_drift(api, orc, sweep_kernel):
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = small_system(np.float64)
    nw, seed = 5, 77
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=0.2, use_drift=False, delay_rank=2)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(2, log_accept=True)
    crowd = api.Crowd(s, nw=nw, delay_rank=2)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=0.2, use_drift=False, seed=seed, use_cuda_graph=False, sweep_kernel=sweep_kernel)
    assert np.array_equal(crowd.vmc_sweep(2, log_accept=True), olog)


def test_mixed_precision_sweep_statistically_equal(api, orc):
    """mixed precision (float tables / inverse): moves agree until the first rounding-borderline decision; energies of
    the ensembles stay statistically equal and delayed updates stay consistent with a from-scratch recompute"""
    s = small_system(np.float32, N=32, M=8)
    crowd, ov, log, olog = host_sweeps(api, orc, s, nw=16, k=4, nsteps=4, tau=0.1)
    assert abs(log.mean() - olog.mean()) < 0.05
    assert (log[0] == olog[0]).mean() > 0.97   # first sweep: essentially every decision identical
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert abs(ke.mean() - oke.mean()) < 4 * (ke.std() + oke.std()) / np.sqrt(len(ke)) + 0.05 * abs(oke.mean())
    crowd.mw_recompute()
    lp2, _, _, _ = crowd.mw_evaluateGL()
    assert lp2 == pytest.approx(lp, rel=2e-4, abs=2e-3)


def _a32_forced(api, orc, dt, tau=0.3, nw=4, seed=31):
    """product sweep (host-driven C-ABI calls) at the NiO-a32 shape, then the oracle teacher-forced with the product's
    accept flags; returns per-move total ratios of both"""
    from qmcpack_b200.workload import make_system, initial_positions
    from qmcpack_b200 import vmc_host
    import oracle_lib
    N = 384
    s = make_system(N=N, M=48, dtype=dt)
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=32)
    crowd.set_positions(R)
    crowd.mw_recompute()
    rng = orc.rng(seed)
    log = np.zeros((1, N, nw), np.uint8)
    ratios = np.zeros((1, N, nw))
    vmc_host.advance_walkers(crowd, rng, tau=tau, log_accept=log[0], log_ratio=ratios[0])
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=32)
    ov.set_positions(R)
    ov.recompute()
    oratios = ov.sweep_forced(log)
    return s, R, crowd, ov, log, ratios, oratios


def _double_twin_ratios(orc, s, R, log, nw, seed, tau, k, use_drift=True):
    """the oracle in FULL precision on the same table values (the float coefficients widened exactly), teacher-forced with
    the same accept flags: the yardstick that tells how far a float32 implementation of this path may legitimately be
    from the exact ratios"""
    import oracle_lib
    sd = dict(s)
    sd["coefs"] = [np.ascontiguousarray(c.astype(np.float64)) for c in s["coefs"]]
    ovd = oracle_lib.OracleVMC(orc, sd, nw=nw, ncrowds=1, seeds=[seed], tau=tau, use_drift=use_drift, delay_rank=k)
    ovd.set_positions(R)
    ovd.recompute()
    return ovd.sweep_forced(log)


def _same_size_as_reference_float(ratios, oratios_f, oratios_d, sel=slice(None), factor=4.0):
    """product(float) vs exact and oracle(float) vs exact must be errors of the same size: median and 95th percentile of
    the product's relative error within `factor` of the float oracle's"""
    den = np.maximum(np.abs(oratios_d), 0.1)
    e_prod = (np.abs(ratios - oratios_d) / den)[0, sel].ravel()
    e_orc = (np.abs(oratios_f - oratios_d) / den)[0, sel].ravel()
    for q in (50, 95):
        assert np.percentile(e_prod, q) <= factor * np.percentile(e_orc, q) + 1e-6, (q, np.percentile(e_prod, q),
                                                                                    np.percentile(e_orc, q))
    return e_prod, e_orc


def test_nio_a32_shape_fp64_every_move(api, orc):
    """BASELINE config 2 shape (384 electrons, 192 orbitals/spin, 48^3 grid, delay rank 32) in full precision: every
    move's total wavefunction ratio within 1e-8 relative of the oracle, and the free-running oracle takes the same
    decisions (identical acceptance sequence)."""
    import oracle_lib
    s, R, crowd, ov, log, ratios, oratios = _a32_forced(api, orc, np.float64)
    rel = np.abs(ratios - oratios) / np.maximum(np.abs(oratios), 1e-2)
    assert rel.max() < 1e-8, rel.max()
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-10, abs=1e-10)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-10, abs=1e-8)
    assert ke == pytest.approx(oke, rel=1e-7)
    ov2 = oracle_lib.OracleVMC(orc, s, nw=4, ncrowds=1, seeds=[31], tau=0.3, delay_rank=32)
    ov2.set_positions(R)
    ov2.recompute()
    assert np.array_equal(ov2.sweep(1, log_accept=True), log)


def test_nio_a32_shape_mixed_precision_every_move(api, orc):
    """Same shape with float tables / float inverse (the benchmark's precision).  A rounding-borderline Metropolis
    decision makes two correct mixed-precision implementations part ways, so the oracle is teacher-forced with the
    product's decisions and EVERY move's ratio is compared (median relative error < 1e-4, 95th percentile < 5e-3,
    worst move < 5e-2: float inverse rows of a Slater matrix with condition number ~1e3).  Then the free-running device driver is checked for the
    oracle's acceptance rate and for consistency with a from-scratch recompute (checkGL_after_moves)."""
    import oracle_lib
    s, R, crowd, ov, log, ratios, oratios = _a32_forced(api, orc, np.float32)
    N, nw = 384, 4
    rel = np.abs(ratios - oratios) / np.maximum(np.abs(oratios), 0.1)
    # float32 inverse rows of a Slater matrix with condition number ~1e3, errors compounding over 384 rank-1 updates
    assert rel.max() < 5e-2, rel.max()
    assert np.percentile(rel, 95) < 5e-3
    assert np.median(rel) < 1e-4
    # anchor of those tolerances: against the FP64 oracle on the same table values the product's float error is the size
    # of the reference float path's own error (a kernel bug at the 1e-3 level, e.g. in the TF32x3 split, would not be)
    oratios_d = _double_twin_ratios(orc, s, R, log, nw, 31, 0.3, 32)
    e_prod, e_orc = _same_size_as_reference_float(ratios, oratios, oratios_d)
    post_flush = slice(32, None)  # every move after the first Woodbury flush of the first determinant
    _same_size_as_reference_float(ratios, oratios, oratios_d, sel=post_flush)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-5, abs=2e-2)
    assert ke == pytest.approx(oke, rel=5e-3)
    ov2 = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[31], tau=0.3, delay_rank=32)
    ov2.set_positions(R)
    ov2.recompute()
    olog = ov2.sweep(1, log_accept=True)
    for sweep_kernel in (1, 2):
        crowd.set_positions(R)
        crowd.mw_recompute()
        crowd.vmc_init(tau=0.3, use_drift=True, seed=31, use_cuda_graph=True, sweep_kernel=sweep_kernel)
        dlog = crowd.vmc_sweep(1, log_accept=True)
        assert abs(dlog.mean() - olog.mean()) < 0.03
        assert (dlog == olog).mean() > 0.9
        lp, ke, _, _ = crowd.mw_evaluateGL()
        crowd.mw_recompute()
        lp2, ke2, _, _ = crowd.mw_evaluateGL()
        assert lp2 == pytest.approx(lp, rel=1e-5, abs=2e-2)
        assert ke2 == pytest.approx(ke, rel=5e-3)


# ---------------------------------------------------------------------------------------------------------------------
# complex orbitals (SplineC2C tables, complex determinants): the NiO-a128 class of BASELINE.json at test size
# ---------------------------------------------------------------------------------------------------------------------
def complex_system(dt, lattice=None, N=24, M=8):
    from qmcpack_b200.workload import make_system
    return make_system(N=N, M=M, dtype=dt, L=6.0, lattice=lattice, complex_orbitals=True)


@pytest.mark.parametrize("lattice", [None, LAT_GENERAL], ids=["ortho", "general"])
@pytest.mark.parametrize("k", [1, 4])
def test_complex_orbitals_host_driven_identical_acceptance_fp64(api, orc, lattice, k):
    """complex determinants, FP64: prob = |ratio|^2 and the real part of the complex gradient drive the same decisions
    as the oracle; complex G, L and the kinetic energy real(G.G + sum L) agree to rounding"""
    s = complex_system(np.float64, lattice)
    crowd, ov, log, olog = host_sweeps(api, orc, s, nw=6, k=k, nsteps=3, tau=0.1)
    assert crowd.cplx and 0.2 < olog.mean() < 0.98
    assert np.array_equal(log, olog)
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)
    lp, ke, G, L = crowd.mw_evaluateGL()
    olp, oke, oG, oL = ov.evaluate_gl()
    assert np.iscomplexobj(G) and np.abs(G.imag).max() > 1e-3  # genuinely complex wavefunction
    assert lp == pytest.approx(olp, rel=1e-8, abs=1e-8)
    assert ke == pytest.approx(oke, rel=1e-6, abs=1e-6)
    assert np.abs(G - oG).max() < 1e-6 * max(1.0, np.abs(oG).max())
    assert np.abs(L - oL).max() < 1e-6 * max(1.0, np.abs(oL).max())
    for spin in (0, 1):
        inv, ld = crowd.det_mw_completeUpdates(spin)
        for iw in range(2):
            oinv, old = ov.psiminv(iw, spin)
            assert np.abs(inv[iw] - oinv).max() < 1e-8 * max(1.0, np.abs(oinv).max())
            assert ld[iw, 0] == pytest.approx(old, rel=1e-9, abs=1e-9)
    crowd.mw_recompute()
    lp2, ke2, _, _ = crowd.mw_evaluateGL()
    assert lp2 == pytest.approx(lp, rel=1e-9, abs=1e-9)
    assert ke2 == pytest.approx(ke, rel=1e-7, abs=1e-7)


def test_complex_orbitals_device_driver_identical_acceptance_fp64(api, orc):
    """device-resident sweep with complex determinants (with and without CUDA graph replay) against the oracle"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = complex_system(np.float64)
    nw, k, nsteps, tau, seed = 7, 4, 3, 0.1, 4243
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    for graph in (False, True):
        crowd = api.Crowd(s, nw=nw, delay_rank=k)
        crowd.set_positions(R)
        crowd.mw_recompute()
        crowd.vmc_init(tau=tau, use_drift=True, seed=seed, use_cuda_graph=graph)
        log = crowd.vmc_sweep(nsteps, log_accept=True)
        assert np.array_equal(log, olog), f"graph={graph}: {np.argwhere(log != olog)[:5]}"
        assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)


def test_complex_orbitals_compiled_host_driver(api, orc):
    """the C++ host driver above the C ABI with complex ratios / gradients (interleaved doubles), two crowds"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = complex_system(np.float64)
    nw, k, nsteps, tau, ncrowds = 6, 4, 2, 0.1, 2
    seeds = [21, 26]
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=ncrowds, seeds=seeds, tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    crowds, spo = [], None
    for c in range(ncrowds):
        cr = api.Crowd(s, nw=3, delay_rank=k, spo=spo)
        spo = cr.spo
        cr.set_positions(R[3 * c:3 * c + 3])
        cr.mw_recompute()
        crowds.append(cr)
    drv = api.HostVMC(crowds, seeds, tau=tau, use_drift=True)
    log = drv.run(nsteps, log_accept=True)
    assert np.array_equal(log, olog)
    got = np.concatenate([c.positions() for c in crowds])
    assert got == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)


def test_complex_orbitals_mixed_precision_every_move(api, orc):
    """complex<float> determinants on float tables: the oracle is teacher-forced with the product's decisions and every
    move's complex ratio is compared; then delayed updates against a from-scratch recompute"""
    from qmcpack_b200.workload import initial_positions
    from qmcpack_b200 import vmc_host
    import oracle_lib
    N, nw, k, tau, seed = 48, 6, 8, 0.2, 91
    s = complex_system(np.float32, N=N, M=10)
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    rng = orc.rng(seed)
    log = np.zeros((2, N, nw), np.uint8)
    ratios = np.zeros((2, N, nw), np.complex128)
    for step in range(2):
        vmc_host.advance_walkers(crowd, rng, tau=tau, log_accept=log[step], log_ratio=ratios[step])
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    oratios = ov.sweep_forced(log)
    rel = np.abs(ratios - oratios) / np.maximum(np.abs(oratios), 0.1)
    assert np.median(rel) < 1e-4 and rel.max() < 5e-2, (np.median(rel), rel.max())
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-4, abs=2e-2)
    assert ke == pytest.approx(oke, rel=2e-2)
    crowd.mw_recompute()
    lp2, _, _, _ = crowd.mw_evaluateGL()
    assert lp2 == pytest.approx(lp, rel=2e-4, abs=5e-3)


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["fp64", "mixed"])
def test_nio_a256_orbital_count_every_move(api, orc, dt):
    """The widest determinant of BASELINE.json (NiO-a256: 1536 orbitals per spin, delay rank 32) on a small grid: 48
    partial-dot slots per walker (more than one warp of them), 8 gather tiles per walker, the tcgen05 flush in 128-column
    pieces.  The oracle is teacher-forced with the product's decisions over the first 96 electrons of each spin and every
    move's ratio is compared."""
    from qmcpack_b200.workload import make_system, initial_positions
    from qmcpack_b200 import vmc_host
    import oracle_lib
    N, nw, k, tau, seed = 3072, 2, 32, 0.3, 77
    s = make_system(N=N, M=20, dtype=dt, with_j1=True, with_j2=True)
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    rng = orc.rng(seed)
    log = np.zeros((1, N, nw), np.uint8)
    ratios = np.zeros((1, N, nw))
    vmc_host.advance_walkers(crowd, rng, tau=tau, use_drift=False, log_accept=log[0], log_ratio=ratios[0])
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, use_drift=False, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    oratios = ov.sweep_forced(log)
    rel = np.abs(ratios - oratios) / np.maximum(np.abs(oratios), 0.1)
    # Uniformly random (unequilibrated) positions of 1536 electrons per spin give Slater matrices whose inverse amplifies
    # rounding differences by ~10x every 200 rank-1 updates (measured: 1e-12 at move 0, 1e-9 at move 64, 1e-6 at move 600,
    # reset at the spin change), in the oracle exactly as in the product.  The moves are made WITHOUT drift so that the
    # positions (R + sqrt(tau) * Gaussian, identical streams) stay bit-identical on both sides and only the ratios carry
    # the amplified rounding; the comparison is tight over the first 256 electrons of each determinant (8 Woodbury
    # flushes each) and statistical over the rest.
    n = N // 2
    early = np.concatenate([rel[0, :256], rel[0, n:n + 256]])
    if dt == np.float64:
        assert early.max() < 1e-7, early.max()
        assert np.median(rel) < 1e-4
        assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-12, abs=1e-12)
        # the device-resident driver at this width (48 partial-dot slots > one warp of them): same stream, same decisions
        # as the host-driven loop over the span where rounding has not been amplified yet
        dev = api.Crowd(s, nw=nw, delay_rank=k, spo=crowd.spo)
        dev.set_positions(R)
        dev.mw_recompute()
        dev.vmc_init(tau=tau, use_drift=False, seed=seed, use_cuda_graph=False)
        dlog = dev.vmc_sweep(1, log_accept=True)
        assert np.array_equal(dlog[0, :512], log[0, :512])
        assert np.array_equal(dlog[0, n:n + 256], log[0, n:n + 256])
        assert abs(dlog.mean() - log.mean()) < 0.02
    else:
        first = np.concatenate([rel[0, :32], rel[0, n:n + 32]])  # before the first flush of each determinant
        assert np.median(first) < 1e-3 and first.max() < 0.2, (np.median(first), first.max())
        # AFTER the flushes (moves 32..255 of each determinant: seven tcgen05 flushes at n = 1536): the product's error
        # against the FP64 oracle is the size of the float oracle's own error against it
        oratios_d = _double_twin_ratios(orc, s, R, log, nw, seed, tau, k, use_drift=False)
        for lo in (0, n):
            _same_size_as_reference_float(ratios, oratios, oratios_d, sel=slice(lo + 32, lo + 256))
    # the product's own delayed-update state against ITS from-scratch recompute (checkGL_after_moves of the reference)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    crowd.mw_recompute()
    lp2, ke2, _, _ = crowd.mw_evaluateGL()
    if dt == np.float64:
        assert lp2 == pytest.approx(lp, rel=1e-9, abs=1e-5)
        assert ke2 == pytest.approx(ke, rel=1e-5)
    else:
        assert np.isfinite(lp).all() and np.isfinite(ke).all()


@pytest.mark.parametrize("cplx,nw,force", [(True, 2048, False), (False, 700, False), (False, 40, True)],
                         ids=["complex_2048_walkers", "real_700_walkers", "forced_small"])
def test_oversubscribed_crowd_look_back_by_ticket(api, orc, cplx, nw, force, monkeypatch):
    """one crowd with more walkers than the device can hold boundary-kernel CTAs at once: the walker index of a CTA is
    its start order (DriverDev::ticket), so the cross-CTA RNG-order wait of the Metropolis test (VMCBatched.cpp:156-158:
    the uniform is drawn only when prob >= eps, walkers in index order) cannot wait on a CTA that has not started.
    Identical acceptance sequence with the oracle in FP64, two sweeps, both with and without CUDA-graph replay."""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    if force:
        monkeypatch.setenv("QMCB_TICKET", "1")
    s = (complex_system if cplx else small_system)(np.float64, None, N=12, M=6)
    k, nsteps, tau, seed = 3, 2, 0.15, 99
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(nsteps, log_accept=True)
    for graph in (False, True):
        crowd = api.Crowd(s, nw=nw, delay_rank=k)
        crowd.set_positions(R)
        crowd.mw_recompute()
        crowd.vmc_init(tau=tau, use_drift=True, seed=seed, use_cuda_graph=graph, sweep_kernel=1)
        assert crowd.sweep_kernel == 1
        log = crowd.vmc_sweep(nsteps, log_accept=True)
        assert np.array_equal(log, olog), f"graph={graph}: {np.argwhere(log != olog)[:5]}"
        assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-8, abs=1e-8)
