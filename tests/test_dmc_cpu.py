"""Host logic of the DMC layer (qmcpack_b200/dmc.py) on CPU: branch weights, multiplicities, the swapWalkersSimple
schedule, and a 2-rank gloo run of WalkerControl::branch with a fake engine whose walkers are tagged byte buffers."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from qmcpack_b200 import dmc, sharding  # noqa: E402


def test_branch_weight_matches_scalar_formula():
    """SFNBranch::branchWeight (SFNBranch.h:199-208) incl. the filtered window between cut-off and 1.5 cut-off"""
    b = dmc.BranchEngine(tau=0.01, target_walkers=100, e_trial=-10.0)
    b.e_ref = -10.2
    b.tau_eff = 0.009
    b.cutoff, b.bmax = 1.0, 1.5
    b.bfilter = 1.0 / (b.bmax - b.cutoff)
    for enew, eold in [(-10.1, -10.3), (-11.5, -10.0), (-12.5, -10.0), (-9.0, -9.5)]:
        taueff = 0.5 * b.tau_eff
        x = max(b.e_ref - enew, b.e_ref - eold)
        if x > b.bmax:
            taueff = 0.0
        elif x > b.cutoff:
            taueff *= 1.0 - (x - b.cutoff) * b.bfilter
        ref = np.exp(taueff * (2 * b.e_trial - enew - eold))
        assert b.branch_weight([enew], [eold])[0] == pytest.approx(ref, rel=1e-14)


def test_multiplicities_truncate_weight_plus_uniform():
    m = dmc.multiplicities([0.3, 0.3, 1.0, 1.7, 2.2], [0.6, 0.8, 0.5, 0.4, 0.9])
    assert list(m) == [0, 1, 1, 2, 3]


@pytest.mark.parametrize("seed", range(6))
def test_swap_schedule_balances_and_conserves_population(seed):
    """every rank derives its part of the schedule alone; replaying all parts must leave exactly the fair share on
    every rank (the Multiplicity check of WalkerControl.cpp:478-493) and never send a dead walker"""
    rng = np.random.default_rng(seed)
    world = int(rng.integers(2, 6))
    mults = [list(rng.integers(0, 4, size=int(rng.integers(3, 9)))) for _ in range(world)]
    for m in mults:
        if sum(m) == 0:
            m[0] = 1
    num_per_rank = [int(sum(m)) for m in mults]
    fair, minus, plus = sharding.determine_new_walker_population(num_per_rank)
    received = [0] * world
    after = []
    for r in range(world):
        sends, m_after = dmc.swap_jobs(r, num_per_rank, mults[r])
        for widx, target, extra in sends:
            assert mults[r][widx] > 0
            received[target] += 1 + extra
        after.append(m_after)
    for r in range(world):
        assert sum(after[r]) + received[r] == fair[r + 1] - fair[r]
    assert sum(sum(a) for a in after) + sum(received) == sum(num_per_rank)


class FakeEngine:
    """walkers are 64-byte tagged buffers held in a torch CPU tensor (so data_ptr() is a real address)"""

    def __init__(self, capacity, tags):
        import torch
        self.capacity = capacity
        self.nw = len(tags)
        self.walker_bytes = 64
        self.state = torch.zeros((capacity, 64), dtype=torch.uint8)
        for i, t in enumerate(tags):
            self.state[i] = t
        self.e = np.array([float(t) for t in tags])

    def dmc_sweep(self):
        pass

    def local_energies(self):
        return np.array([float(self.state[i, 0]) for i in range(self.nw)])

    def rr(self):
        return np.full(self.nw, 0.5), np.ones(self.nw)

    def copy_walker(self, src, dst):
        if src != dst:
            self.state[dst] = self.state[src].clone()

    def set_num_walkers(self, n):
        self.nw = n

    def pack_walker(self, iw, ptr):
        import ctypes
        ctypes.memmove(ptr, self.state[iw].data_ptr(), 64)

    def unpack_walker(self, iw, ptr):
        import ctypes
        ctypes.memmove(self.state[iw].data_ptr(), ptr, 64)


def _branch_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 owns heavy walkers, rank 1 light ones: after branching rank 0 must ship walkers to rank 1
    tags = [10, 11, 12, 13] if rank == 0 else [20, 21, 22, 23]
    eng = FakeEngine(16, tags)
    d = dmc.DMC(eng, tau=0.01, target_walkers=8, branch_rng=lambda: 0.5, dist=dist)
    d.weights = np.array([3.0, 2.0, 1.0, 2.6]) if rank == 0 else np.array([0.2, 1.0, 0.1, 0.4])
    ens = d.branch_step()
    q.put((rank, eng.nw, [int(eng.state[i, 0]) for i in range(eng.nw)], list(d.energies), ens["weight"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_branch_moves_walkers_with_their_state():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_branch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, tags0, e0, w0), (r1, n1, tags1, e1, w1) = out
    # multiplicities: rank 0 -> int(w + .5) = [3, 2, 1, 3] = 9 copies; rank 1 -> [0, 1, 0, 0] = 1 copy; fair share 5 + 5
    assert (n0, n1) == (5, 5)
    assert w0 == pytest.approx(10.3) and w1 == pytest.approx(10.3)
    # rank 1 keeps its survivor first, then holds copies of rank 0's walkers; energies travel with the walkers
    assert tags1[0] == 21 and all(t in (10, 11, 12, 13) for t in tags1[1:])
    assert all(t in (10, 11, 12, 13) for t in tags0)
    assert sorted(tags0 + tags1[1:]) == sorted([10] * 3 + [11] * 2 + [12] + [13] * 3)
    assert e0 == [float(t) for t in tags0] and e1 == [float(t) for t in tags1]
