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


# ---------------------------------------------------------------------------------------------------------------------
# the C++ DMC layer (csrc/dmc_host.cpp, bound as api.DMCDriver) against this Python mirror and against hand arithmetic
# ---------------------------------------------------------------------------------------------------------------------
class NoisyEngine:
    """walkers are rows of a numpy array: column 0 a tag, column 1 the last energy fluctuation"""

    def __init__(self, capacity, tags, spread=1.0):
        self.capacity, self.nw, self.spread = capacity, len(tags), spread
        self.state = np.zeros((capacity, 8))
        self.state[:len(tags), 0] = tags
        self.t = 0

    def dmc_sweep(self):
        self.t += 1
        self.state[:self.nw, 1] = self.spread * np.random.default_rng(self.t).normal(size=self.nw)

    def local_energies(self):
        return -10 + 0.01 * self.state[:self.nw, 0] + self.state[:self.nw, 1]

    def rr(self):
        f = 1 + self.state[:self.nw, 0] % 3
        return 0.5 * f, 1.0 * f

    def copy_walker(self, src, dst):
        self.state[dst] = self.state[src]

    def set_num_walkers(self, n):
        self.nw = n


def _have_lib():
    from qmcpack_b200 import api
    return os.path.exists(api.LIB_PATH)


pytestmark_lib = pytest.mark.skipif(not _have_lib(), reason="libqmcb.so not built")


@pytestmark_lib
@pytest.mark.parametrize("spread", [0.3, 6.0])
def test_cpp_dmc_layer_matches_python_mirror(orc, spread):
    """same engine, same branching stream: populations, walker order, weights, ages, E_trial and tau_eff identical over 25
    generations of a fluctuating population (warm-up stage; iteration 0 does not branch and keeps its weights)"""
    from qmcpack_b200 import api
    tags = list(range(10, 22))
    ea, eb = NoisyEngine(256, tags, spread), NoisyEngine(256, tags, spread)
    da = dmc.DMC(ea, 0.05, 12, orc.rng(7).uniform)
    db = api.DMCDriver(eb, 0.05, 12, branch_seed=7)
    pops = []
    for gen in range(25):
        da.advance()
        db.advance()
        wb, enb, ageb = db.walkers()
        assert da.weights == pytest.approx(wb, rel=1e-13)
        assert np.array_equal(da.ages, ageb)
        xa = da.branch_step(do_not_branch=(gen == 0))
        xb = db.branch_step(gen, do_not_branch=(gen == 0))
        assert ea.nw == eb.nw == xb["local"] == xb["population"]
        assert np.array_equal(ea.state[:ea.nw], eb.state[:eb.nw])  # same walkers in the same order
        wb, enb, ageb = db.walkers()
        assert da.weights == pytest.approx(wb, rel=1e-13) and da.energies == pytest.approx(enb, rel=1e-13)
        if gen == 0:
            assert not np.allclose(wb, 1.0)  # WalkerControl.cpp:226: no reset when do_not_branch
        else:
            assert np.all(wb == 1.0)
        for key in ("energy", "variance", "weight", "r2_accepted", "r2_proposed", "living_fraction"):
            assert xa[key] == pytest.approx(xb[key], rel=1e-12, abs=1e-12)
        assert da.branch.e_trial == pytest.approx(xb["e_trial"], rel=1e-12)
        assert da.branch.tau_eff == pytest.approx(xb["tau_eff"], rel=1e-12)
        pops.append(eb.nw)
    if spread > 1:
        assert len(set(pops)) > 3, pops


@pytestmark_lib
def test_cpp_dmc_main_stage_rules():
    """SFNBranch::updateParamAfterPopControl after warm-up (SFNBranch.cpp:150-173): E_ref = running mean of the
    generation energies since the end of warm-up, E_trial = E_ref + feedback (log N_target - log W) every
    `energy_update_interval` generations; the cut-off is reset from the variance at the end of warm-up (:186-187)"""
    from qmcpack_b200 import api
    eng = NoisyEngine(256, list(range(10, 30)), spread=0.5)
    tau, target, fb = 0.02, 20, 0.7
    d = api.DMCDriver(eng, tau, target, branch_seed=3, warmup_steps=3, energy_update_interval=2, feedback=fb,
                      sigma_bound=10.0)
    hist, e_trial, todo = [], None, 2
    for gen in range(9):
        x = d.step(gen)
        if gen < 3:  # warm-up
            assert x["e_ref"] == pytest.approx(x["energy"])
            assert x["e_trial"] == pytest.approx(x["energy"] + (np.log(target) - np.log(x["weight"])) / tau, rel=1e-12)
            e_trial = x["e_trial"]
            if gen == 2:
                assert x["branch_cutoff"] == pytest.approx(min(max(x["variance"] * 10.0, 10.0), 2.5 / tau))
        else:
            hist.append(x["energy"])
            assert x["e_ref"] == pytest.approx(np.mean(hist), rel=1e-12)
            todo -= 1
            if todo == 0:
                e_trial = np.mean(hist) + fb * (np.log(target) - np.log(x["weight"]))
                todo = 2
            assert x["e_trial"] == pytest.approx(e_trial, rel=1e-12)


@pytestmark_lib
def test_cpp_dmc_capacity_overflow_is_an_error():
    from qmcpack_b200 import api
    eng = NoisyEngine(6, [1, 2, 3, 4])
    d = api.DMCDriver(eng, 0.05, 4, branch_seed=1)
    d.set_weights([3.0, 3.0, 3.0, 3.0])
    with pytest.raises(RuntimeError, match="capacity"):
        d.branch_step(1)


def _cpp_branch_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from qmcpack_b200 import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tags = [10, 11, 12, 13] if rank == 0 else [20, 21, 22, 23]
    eng = FakeEngine(16, tags)
    eng.rr = lambda: (np.full(eng.nw, 0.5), np.ones(eng.nw))
    comm = api.TorchComm(dist, eng.walker_bytes, "cpu")
    d = api.DMCDriver(eng, 0.01, 8, branch_seed=5, comm=comm)
    d.set_weights([3.0, 2.0, 1.0, 2.6] if rank == 0 else [0.2, 1.0, 0.1, 0.4])
    # the multiplicities int(w + u) with the layer's own std::mt19937(5) stream
    import oracle_lib
    u = oracle_lib.port().rng(5)
    w0 = np.array([3.0, 2.0, 1.0, 2.6] if rank == 0 else [0.2, 1.0, 0.1, 0.4])
    mult = [int(x + u.uniform()) for x in w0]
    ens = d.branch_step(1)
    w, e, a = d.walkers()
    q.put((rank, eng.nw, [int(eng.state[i, 0]) for i in range(eng.nw)], list(e), ens["weight"], mult, ens["population"],
           int(ens["walkers_sent"]), int(ens["walkers_received"]), list(w)))
    dist.barrier()
    dist.destroy_process_group()


@pytestmark_lib
def test_cpp_dmc_two_rank_branch_moves_walkers_with_their_state():
    """WalkerControl::branch + swapWalkersSimple of the C++ layer across two gloo ranks: the global population is the sum
    of the multiplicities, each rank ends with its fair share, walkers arrive with their energy and state"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_cpp_branch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, tags0, e0, w0, m0, pop0, s0, rcv0, ww0), (r1, n1, tags1, e1, w1, m1, pop1, s1, rcv1, ww1) = out
    total = sum(m0) + sum(m1)
    assert pop0 == pop1 == total and n0 + n1 == total and abs(n0 - n1) <= 1
    assert w0 == pytest.approx(10.3) and w1 == pytest.approx(10.3)
    assert s0 + s1 >= 1 and s0 + s1 == (rcv0 + rcv1)  # messages (each may carry folded copies)
    # copies present after branching = multiplicity of each tag
    want = sorted(sum(([t] * m for t, m in zip([10, 11, 12, 13], m0)), []) + sum(([t] * m for t, m in zip([20, 21, 22, 23], m1)), []))
    assert sorted(tags0 + tags1) == want
    assert e0 == [float(t) for t in tags0] and e1 == [float(t) for t in tags1]
    assert all(x == 1.0 for x in ww0 + ww1)
