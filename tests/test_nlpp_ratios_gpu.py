"""the ratio path of the non-local pseudopotential and the ratio-only move.

  TrialWaveFunction::mw_evaluateRatios   TrialWaveFunction.cpp:1079-1110 (determinant rows of psiMinv x V-only orbitals at the
                                         quadrature points; TwoBodyJastrow::mw_evaluateRatios; J1 evaluateRatios)
  TrialWaveFunction::mw_calcRatio        TrialWaveFunction.cpp:494-510 (no-drift sweeps)
against the oracle's restatement (oracle/qmc_oracle_driver.hpp evaluateRatios, advanceCrowd without drift)."""
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


@pytest.mark.parametrize("cplx", [False, True], ids=["R2R", "C2C"])
@pytest.mark.parametrize("lattice", [None, LAT_GENERAL], ids=["ortho", "general"])
def test_evaluate_ratios_at_virtual_positions(api, orc, cplx, lattice):
    """quadrature-like virtual positions around several electrons of several walkers, after a sweep that leaves delayed
    updates pending: all three compute types against the oracle; ALL = FERMIONIC x NONFERMIONIC"""
    from qmcpack_b200.workload import make_system, initial_positions
    from qmcpack_b200 import vmc_host
    import oracle_lib
    s = make_system(N=24, M=8, dtype=np.float64, L=6.0, lattice=lattice, complex_orbitals=cplx)
    nw, k, seed = 5, 4, 17
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=0.1, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    # one sweep on both sides (identical streams) so that the inverse is a delayed-update product, not a fresh inverse
    log = np.zeros((24, nw), np.uint8)
    vmc_host.advance_walkers(crowd, orc.rng(seed), tau=0.1, log_accept=log)
    olog = ov.sweep(1, log_accept=True)
    assert np.array_equal(log, olog[0])
    Rn = crowd.positions()
    rng = np.random.default_rng(4)
    walker, ref, rvp = [], [], []
    for iw in (0, 2, 4, 1):
        for iat in (3, 17):  # one electron of each spin
            for q in range(6):  # "quadrature points" on a sphere around an ion-like centre near the electron
                u = rng.normal(size=3)
                u /= np.linalg.norm(u)
                walker.append(iw)
                ref.append(iat)
                rvp.append(Rn[iw, iat] + 0.8 * u)
    walker, ref, rvp = np.array(walker), np.array(ref), np.array(rvp)
    got = {ct: crowd.mw_evaluateRatios(walker, ref, rvp, ct) for ct in (0, 1, 2)}
    for ct in (0, 1, 2):
        want = np.array([ov.evaluate_ratios(int(walker[i]), int(ref[i]), rvp[i:i + 1], ct)[0] for i in range(len(walker))])
        if not cplx:
            want = want.real
        assert got[ct] == pytest.approx(want, rel=1e-9, abs=1e-12), ct
    assert got[0] == pytest.approx(got[1] * got[2], rel=1e-12)
    if cplx:
        assert np.abs(got[1].imag).max() > 1e-6 and np.abs(got[2].imag).max() == 0.0
    # empty request and a bad index
    assert crowd.mw_evaluateRatios([], [], np.zeros((0, 3))).shape == (0,)
    with pytest.raises(RuntimeError, match="out of range"):
        crowd.mw_evaluateRatios([nw], [0], np.zeros((1, 3)))


@pytest.mark.parametrize("k", [1, 4])
def test_calc_ratio_only_sweep_matches_oracle_without_drift(api, orc, k):
    """a no-drift sweep driven with mw_calcRatio (values only): ratios equal those of mw_calcRatioGrad, the acceptance
    sequence equals the oracle's no-drift sweep, and G, L, the kinetic energy afterwards (gradient / Laplacian rows
    re-evaluated because the accepts left them stale) equal the oracle's"""
    from qmcpack_b200.workload import make_system, initial_positions
    import oracle_lib
    s = make_system(N=24, M=8, dtype=np.float64, L=6.0)
    nw, seed, tau = 6, 23, 0.2
    R = initial_positions(s, nw)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, use_drift=False, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    olog = ov.sweep(2, log_accept=True)
    rng = orc.rng(seed)
    eps = np.finfo(np.float64).eps
    for step in range(2):
        deltas = rng.gauss(3 * nw * 24, np.float64).reshape(24, nw, 3) * np.sqrt(tau)
        for iat in range(24):
            crowd.mw_makeMove(iat, deltas[iat])
            r = crowd.mw_calcRatio(iat)
            if iat in (0, 13):  # same proposal through the gradient path gives the same ratio
                r2, _ = crowd.mw_calcRatioGrad(iat)
                assert r == pytest.approx(r2, rel=1e-12)
                r = crowd.mw_calcRatio(iat)
            acc = np.zeros(nw, np.uint8)
            for iw in range(nw):
                if r[iw] * r[iw] >= eps and rng.uniform() < r[iw] * r[iw]:
                    acc[iw] = 1
            assert np.array_equal(acc, olog[step, iat]), (step, iat)
            crowd.mw_accept_rejectMove(iat, acc, True)
        crowd.mw_completeUpdates()
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-10, abs=1e-10)
    lp, ke, G, L = crowd.mw_evaluateGL()
    olp, oke, oG, oL = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-9, abs=1e-9)
    assert ke == pytest.approx(oke, rel=1e-7, abs=1e-7)
    assert G == pytest.approx(oG, rel=1e-7, abs=1e-7)
    assert L == pytest.approx(oL, rel=1e-6, abs=1e-6)
    # and the drift path works again afterwards (stale rows were refreshed)
    g = crowd.mw_evalGrad(5)
    crowd.mw_recompute()
    g2 = crowd.mw_evalGrad(5)
    assert g == pytest.approx(g2, rel=1e-8, abs=1e-8)
