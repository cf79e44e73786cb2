"""N > 1 host logic on CPU: world_size-2 gloo processes (the GPU box runs the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fair_divide_rules():
    from qmcpack_b200 import sharding as sh
    assert sh.fair_divide(10, 4) == [3, 3, 2, 2]             # crowds: extra walkers first (FairDivide.h:77-85)
    assert sh.fair_divide_low(10, 4) == [0, 2, 4, 7, 10]     # ranks: extra walkers last (FairDivide.h:129-145)
    assert sh.fair_divide_low(8, 4) == [0, 2, 4, 6, 8]


def test_determine_new_walker_population():
    """WalkerControl.cpp:284-313: surplus ranks appear in `plus`, deficit ranks in `minus`, counts match"""
    from qmcpack_b200 import sharding as sh
    off, minus, plus = sh.determine_new_walker_population([5, 1, 3, 3])
    assert off == [0, 3, 6, 9, 12]
    assert plus == [0, 0] and minus == [1, 1]
    off, minus, plus = sh.determine_new_walker_population([4, 4])
    assert plus == [] and minus == []
    off, minus, plus = sh.determine_new_walker_population([7, 0, 0])
    assert sorted(plus) == [0] * 5 and sorted(minus) == [1, 1, 2, 2, 2]
    assert sh.exchange_plan([3, 1]) == [(0, 1)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from qmcpack_b200 import sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # block estimator: every rank contributes its walkers' sums
        e = np.array([1.0 + rank, 2.0 * (rank + 1), 4.0, 10.0 * rank, 3.0])
        red = sh.reduce_block_estimator(e, dist)
        tmax = sh.max_over_ranks(0.5 + rank, dist)
        # DMC-style rebalancing: rank 0 holds 3 walkers, rank 1 holds 1 -> one walker migrates 0 -> 1
        counts = [3, 1]
        mine = [torch.full((6,), 100.0 * rank + i) for i in range(counts[rank])]
        new = sh.migrate_walkers(mine, counts, rank, dist)
        q.put((rank, red.tolist(), tmax, [t.tolist() for t in new], sh.rank_seed(1000, rank)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduction_and_walker_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, red0, t0, w0, s0), (r1, red1, t1, w1, s1) = out
    assert red0 == red1 == [3.0, 6.0, 8.0, 10.0, 6.0]
    assert t0 == t1 == 1.5
    assert len(w0) == 2 and len(w1) == 2                      # balanced 2 + 2
    assert w1[-1] == [2.0] * 6                                # rank 0's last walker arrived on rank 1
    assert s0 != s1
