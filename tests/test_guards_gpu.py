"""The reference's numerical guards on the device path: a NaN one-particle ratio throws (NaNguard::checkOneParticleRatio,
TrialWaveFunction.cpp:473,508,549) and so does an accepted move whose determinant ratio is zero
(DiracDeterminantBatched.cpp:494-500).  Errors arrive through the ABI's error convention (non-zero return + message)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


def small_crowd(api, nw=4, k=4):
    from qmcpack_b200.workload import make_system, initial_positions
    s = make_system(N=24, M=8, dtype=np.float64, L=6.0)
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(initial_positions(s, nw))
    crowd.mw_recompute()
    return s, crowd


@pytest.mark.parametrize("host_kernel", ["resident", "launch"])
def test_nan_ratio_throws_in_calc_ratio_grad(api, host_kernel, monkeypatch):
    monkeypatch.setenv("QMCB_HOST_KERNEL", host_kernel)
    s, crowd = small_crowd(api)
    crowd.mw_evalGrad(3)
    displ = np.zeros((4, 3))
    displ[2, 1] = np.nan
    crowd.mw_makeMove(3, displ)
    with pytest.raises(RuntimeError, match="NaNguard::checkOneParticleRatio"):
        crowd.mw_calcRatioGrad(3)


def test_zero_ratio_accept_throws_component_level(api):
    """FakeSPO-style: an orbital row of zeros gives ratio 0; accepting it is the reference's 'Report a bug' exception"""
    from test_det_gpu import tiny_system
    n, nw = 6, 2
    crowd = api.Crowd(tiny_system(n, np.float64), nw=nw, delay_rank=2)
    rng = np.random.default_rng(1)
    psiM = 2 * np.eye(n) + 0.1 * rng.normal(size=(nw, n, n))
    crowd.det_recompute_from_matrices(0, psiM)
    phi = np.zeros((5, nw, n))
    phi[0, 1] = rng.normal(size=n)  # walker 0 keeps a zero row, walker 1 a regular one
    crowd.det_set_phi_vgl(0, phi)
    ratios, _ = crowd.det_mw_ratioGrad(0, 2, from_phi=True)
    assert ratios[0] == 0.0 and ratios[1] != 0.0
    with pytest.raises(RuntimeError, match="curRatio is 0"):
        crowd.det_mw_accept_rejectRow(0, 2, [1, 1])


def test_zero_ratio_rejected_is_fine(api):
    from test_det_gpu import tiny_system
    n, nw = 6, 2
    crowd = api.Crowd(tiny_system(n, np.float64), nw=nw, delay_rank=2)
    psiM = 2 * np.eye(n) + 0.1 * np.random.default_rng(1).normal(size=(nw, n, n))
    crowd.det_recompute_from_matrices(0, psiM)
    crowd.det_set_phi_vgl(0, np.zeros((5, nw, n)))
    crowd.det_mw_ratioGrad(0, 2, from_phi=True)
    crowd.det_mw_accept_rejectRow(0, 2, [0, 0])  # pseudo-accepts: no exception
    inv, _ = crowd.det_mw_completeUpdates(0)
    assert np.isfinite(inv).all()


@pytest.mark.parametrize("sweep_kernel", [1, 2], ids=["two_kernel", "segment_kernel"])
def test_nan_position_throws_in_device_sweep(api, sweep_kernel):
    """a walker whose configuration is NaN makes every ratio NaN: the device-resident sweep reports it at its end"""
    s, crowd = small_crowd(api, nw=3)
    R = crowd.positions()
    R[1, 5, 0] = np.nan
    crowd.set_positions(R)
    crowd.vmc_init(tau=0.1, seed=3, use_cuda_graph=False, sweep_kernel=sweep_kernel)
    with pytest.raises(RuntimeError, match="NaNguard::checkOneParticleRatio"):
        crowd.vmc_sweep(1)
