"""Host-side multi-rank logic of the walker-sharded path (one process per GPU).

Walkers never interact during a sweep, so there is NO data-path collective: every rank owns `walkers` walkers and a
replica of the spline tables.  What crosses ranks:

  * the block estimator reduction       Estimators/EstimatorManagerNew.cpp:338 (reduce_in_place_n), :363 (allreduce)
  * the max-over-ranks of the timed region (bench contract)
  * (DMC, next row 8f-3) the walker-count balancing plan of WalkerControl::determineNewWalkerPopulation
    (QMCDrivers/DMC/WalkerControl.cpp:284-313) built on FairDivideLow (Utilities/FairDivide.h:129-145)

Everything here works on any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
import numpy as np


def fair_divide(ntot, npart):
    """Utilities/FairDivide.h:77-85: the first (ntot % npart) partitions get one extra element (crowds of a rank)."""
    bat, residue = divmod(ntot, npart)
    return [bat + 1] * residue + [bat] * (npart - residue)


def fair_divide_low(ntot, npart):
    """Utilities/FairDivide.h:129-145: offsets, the LAST (ntot % npart) partitions get the extra element."""
    bat = ntot // npart
    residue = npart - ntot % npart
    off = [0]
    for i in range(npart):
        off.append(off[-1] + (bat if i < residue else bat + 1))
    return off


def determine_new_walker_population(num_per_rank):
    """WalkerControl.cpp:284-313: (fair_offset, minus, plus); `plus` lists a sender rank once per surplus walker,
    `minus` a receiver rank once per missing walker; len(plus) == len(minus)."""
    fair_offset = fair_divide_low(int(sum(num_per_rank)), len(num_per_rank))
    minus, plus = [], []
    for ip, n in enumerate(num_per_rank):
        dn = n - (fair_offset[ip + 1] - fair_offset[ip])
        if dn > 0:
            plus += [ip] * dn
        elif dn < 0:
            minus += [ip] * (-dn)
    if len(plus) != len(minus):
        raise RuntimeError("walker send/recv pattern does not match")
    return fair_offset, minus, plus


def exchange_plan(num_per_rank):
    """pairs (sender, receiver), one per migrating walker, in the order swapWalkersSimple walks its queues
    (WalkerControl.cpp:312-500)"""
    _, minus, plus = determine_new_walker_population(num_per_rank)
    return list(zip(plus, minus))


def rank_seed(base_seed, rank, crowd=0, crowds_per_rank=1):
    """one std::mt19937 stream per crowd, distinct across ranks (the reference hands out primes per rank/thread,
    Utilities/RandomNumberControl.cpp; any injective map keeps the streams distinct)"""
    return int(base_seed + rank * crowds_per_rank + crowd)


def reduce_block_estimator(values, dist=None, device=None):
    """sum over ranks of the block accumulators {sum E, sum E^2, sum w, n_accept, n_reject, ...} (<= 1 KB, once per block)"""
    import torch
    t = torch.as_tensor(np.asarray(values, np.float64), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


def max_over_ranks(x, dist=None, device=None):
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def migrate_walkers(states, num_per_rank, rank, dist):
    """Moves packed walker states between ranks according to exchange_plan (point-to-point send/recv: ncclSend/ncclRecv on
    the GPU box, gloo here).  `states` is this rank's list of 1-D tensors (one packed walker each); returns the new list.
    Senders give away their LAST walkers (the reference sends the walkers with most copies first; with unit multiplicity
    that is the tail of the list)."""
    import torch
    plan = exchange_plan(num_per_rank)
    states = list(states)
    for sender, receiver in plan:
        if rank == sender:
            dist.send(states.pop(), dst=receiver)
        elif rank == receiver:
            buf = torch.empty_like(states[0]) if states else None
            if buf is None:
                raise RuntimeError("receiver needs at least one walker to know the packed size")
            dist.recv(buf, src=sender)
            states.append(buf)
    return states
