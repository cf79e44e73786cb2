"""GPU parity of the DMC layer: the device move loop of DMCBatched::advanceWalkers (phase rejection, rr accumulators),
walker duplication / packed transfer, and a branching run whose population follows the oracle's step for step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


def small_system(dt=np.float64, complex_orbitals=False):
    from qmcpack_b200.workload import make_system
    return make_system(N=24, M=8, dtype=dt, L=6.0, complex_orbitals=complex_orbitals)


class OracleEngine:
    """adapter: the oracle's VMC object behind the engine interface of qmcpack_b200.dmc.DMC"""

    def __init__(self, ov, capacity):
        self.ov, self.capacity = ov, capacity

    nw = property(lambda self: self.ov.nw)

    def dmc_sweep(self):
        self.last_log = self.ov.sweep(1, log_accept=True)

    def local_energies(self):
        return self.ov.evaluate_gl()[1]

    def rr(self):
        return self.ov.rr()

    def copy_walker(self, src, dst):
        self.ov.copy_walker(src, dst)

    def set_num_walkers(self, n):
        self.ov.set_num_walkers(n)


@pytest.mark.parametrize("cplx", [False, True], ids=["real", "complex"])
def test_dmc_move_loop_identical_acceptance_fp64(api, orc, cplx):
    """node crossings are rejected for real wavefunctions (and only there), prob is tested as a whole, rr sums agree"""
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = small_system(np.float64, cplx)
    nw, k, nsteps, tau, seed = 9, 4, 3, 0.3, 515
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k, dmc=True)
    ov.set_positions(R)
    ov.recompute()
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=tau, use_drift=True, seed=seed, use_cuda_graph=False, dmc=True)
    for step in range(nsteps):
        olog = ov.sweep(1, log_accept=True)
        log = crowd.vmc_sweep(1, log_accept=True)
        assert np.array_equal(log, olog), f"step {step}: {np.argwhere(log != olog)[:5]}"
        a, p = crowd.dmc_rr()
        oa, op = ov.rr()
        assert a == pytest.approx(oa, rel=1e-10) and p == pytest.approx(op, rel=1e-10)
        assert (p >= a).all() and p.min() > 0
    assert crowd.positions() == pytest.approx(ov.positions(), rel=1e-9, abs=1e-9)


def test_walker_copy_and_packed_transfer(api, orc):
    """a duplicated walker and a walker shipped as a packed device buffer into ANOTHER crowd carry the complete state:
    the sweeps that follow match the oracle's, whose walkers were duplicated the same way"""
    import torch
    from qmcpack_b200.workload import initial_positions
    import oracle_lib
    s = small_system()
    nw, k, tau, seed = 6, 4, 0.2, 99
    R = initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
    ov.set_positions(R)
    ov.recompute()
    crowd = api.Crowd(s, nw=nw, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=tau, seed=seed, use_cuda_graph=False)
    assert np.array_equal(crowd.vmc_sweep(1, log_accept=True), ov.sweep(1, log_accept=True))
    # duplicate walker 1 over walker 4, then ship walker 2 through a second crowd and back into slot 0
    crowd.copy_walker(1, 4)
    ov.copy_walker(1, 4)
    other = api.Crowd(s, nw=2, delay_rank=k, spo=crowd.spo)
    other.set_positions(R[:2])
    other.mw_recompute()
    buf = torch.empty(crowd.walker_bytes, dtype=torch.uint8, device="cuda")
    assert other.walker_bytes == crowd.walker_bytes
    crowd.pack_walker(2, buf.data_ptr())
    other.unpack_walker(1, buf.data_ptr())
    buf2 = torch.empty_like(buf)
    other.pack_walker(1, buf2.data_ptr())
    assert torch.equal(buf, buf2)
    crowd.unpack_walker(0, buf2.data_ptr())
    ov.copy_walker(2, 0)
    log, olog = crowd.vmc_sweep(2, log_accept=True), ov.sweep(2, log_accept=True)
    assert np.array_equal(log, olog)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    olp, oke, _, _ = ov.evaluate_gl()
    assert lp == pytest.approx(olp, rel=1e-9, abs=1e-9) and ke == pytest.approx(oke, rel=1e-7)
    # the live count can shrink and grow inside the capacity
    crowd.set_num_walkers(4)
    ov.set_num_walkers(4)
    assert np.array_equal(crowd.vmc_sweep(1, log_accept=True), ov.sweep(1, log_accept=True))
    crowd.set_num_walkers(6)
    ov.set_num_walkers(6)
    assert np.array_equal(crowd.vmc_sweep(1, log_accept=True), ov.sweep(1, log_accept=True))


def test_dmc_branching_run_follows_oracle(api, orc):
    """eight DMC generations with branching (dynamic population inside a capacity of 16): the product -- GPU crowd steered
    by the C++ DMC layer (csrc/dmc_host.cpp) -- and the checker -- the oracle's CPU walkers steered by the Python mirror
    (qmcpack_b200/dmc.py) -- share one branching stream and keep identical populations, acceptance logs, weights and
    energies"""
    from qmcpack_b200.workload import initial_positions
    from qmcpack_b200 import dmc
    import oracle_lib
    s = small_system()
    cap, n0, k, tau, seed = 16, 8, 4, 0.05, 2024
    R = initial_positions(s, cap)
    ov = oracle_lib.OracleVMC(orc, s, nw=cap, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k, dmc=True)
    ov.set_positions(R)
    ov.recompute()
    ov.set_num_walkers(n0)
    crowd = api.Crowd(s, nw=cap, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=tau, seed=seed, use_cuda_graph=True, dmc=True)
    crowd.set_num_walkers(n0)
    da = api.DMCDriver(crowd, tau, n0, branch_seed=7)
    db = dmc.DMC(OracleEngine(ov, cap), tau, n0, orc.rng(7).uniform)
    pops = []
    for gen in range(8):
        da.advance()
        db.advance()
        wa, ena, agea = da.walkers()
        assert wa == pytest.approx(db.weights, rel=1e-7)
        assert ena == pytest.approx(db.energies, rel=1e-7)
        assert np.array_equal(agea, db.ages)
        ea = da.branch_step(gen, do_not_branch=(gen == 0))
        eb = db.branch_step(do_not_branch=(gen == 0))
        assert ea["local"] == ov.nw == ea["population"]
        assert ea["energy"] == pytest.approx(eb["energy"], rel=1e-7)
        assert ea["e_trial"] == pytest.approx(db.branch.e_trial, rel=1e-7, abs=1e-7)
        assert crowd.positions() == pytest.approx(ov.positions()[:crowd.nw], rel=1e-9, abs=1e-9), f"generation {gen}"
        pops.append(crowd.nw)
    assert len(set(pops)) > 1, pops  # the population really changed
    assert crowd.positions() == pytest.approx(ov.positions()[:crowd.nw], rel=1e-9, abs=1e-9)


def _two_gpu_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qmcpack_b200 import api, build
    from qmcpack_b200.workload import initial_positions
    build.build()
    api.init(rank)
    s = small_system()
    cap, n0, k, tau = 24, 8, 4, 0.05
    R = initial_positions(s, cap, seed=7 + 1000 * rank)
    crowd = api.Crowd(s, nw=cap, delay_rank=k)
    crowd.set_positions(R)
    crowd.mw_recompute()
    crowd.vmc_init(tau=tau, seed=100 + rank, use_cuda_graph=False, dmc=True)
    crowd.set_num_walkers(n0)
    comm = api.TorchComm(dist, crowd.walker_bytes, torch.device("cuda", rank))
    d = api.DMCDriver(crowd, tau, world * n0, branch_seed=5 + rank, comm=comm)
    # make rank 0 heavy and rank 1 light so that walkers must cross NVLink
    pops, sent = [], 0
    for gen in range(5):
        d.advance()
        if gen == 1:
            w, _, _ = d.walkers()
            d.set_weights(w * (2.5 if rank == 0 else 0.3))
        ens = d.branch_step(gen, do_not_branch=(gen == 0))
        sent += ens["walkers_sent"]
        pops.append(crowd.nw)
    print("rank %d: populations %s, messages sent %d, %d bytes over NCCL p2p" % (rank, pops, sent, comm.bytes_sent), flush=True)
    lp, ke, _, _ = crowd.mw_evaluateGL()
    crowd.mw_recompute()  # delayed-update state of received / copied walkers == from-scratch state
    lp2, ke2, _, _ = crowd.mw_evaluateGL()
    q.put((rank, pops, bool(np.isfinite(ke).all()), float(np.abs(lp - lp2).max()), float(np.abs(ke - ke2).max()),
           [h["population"] for h in d.history]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_walker_exchange_over_nccl(api):
    """WalkerControl::branch across two GPUs: the per-rank populations stay within one walker of each other after every
    branch, the global population is what the all-reduced multiplicities say, and walkers that crossed NVLink as packed
    device buffers carry a state identical to a from-scratch recompute"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, 29741, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, pops0, fin0, dlp0, dke0, glob0), (_, pops1, fin1, dlp1, dke1, glob1) = out
    assert fin0 and fin1
    assert glob0 == glob1
    for a, b, g in zip(pops0, pops1, glob0):
        assert abs(a - b) <= 1 and a + b == min(g, 48)
    assert pops0[2] != 8 or pops1[2] != 8  # the imbalance of generation 1 really moved walkers
    assert max(dlp0, dlp1) < 1e-8 and max(dke0, dke1) < 1e-6
