"""Python MIRROR of the DMC layer (the product is C++: csrc/dmc_host.cpp behind include/qmcb_driver.h, bound as
qmcpack_b200.api.DMCDriver).  The tests drive the oracle with this file and the GPU crowd with the C++ layer, so the two
restatements check each other.  Branch weights, population control and walker exchange, restating

  * DMCBatched::advanceWalkers, the part after the move loop   QMCDrivers/DMC/DMCBatched.cpp:264-292
  * SFNBranch::branchWeight / setBranchCutoff / warm-up E_trial QMCDrivers/SFNBranch.h:199-208, SFNBranch.cpp:133-199,290-320
  * WalkerControl::branch (dynamic population)                  QMCDrivers/DMC/WalkerControl.cpp:151-240
  * WalkerControl::computeCurData                               :243-281
  * WalkerControl::swapWalkersSimple (who sends what to whom)   :312-500
  * MCPopulation::killWalker / spawnWalker / fissionHighMultiplicityWalkers

The move loop itself (phase rejection, rr accumulators) runs on the device: qmcb_vmc_init(dmc = 1) + qmcb_vmc_sweep.  A
walker's state never visits the host: copies inside a rank are device-to-device (qmcb_crowd_copy_walker), walkers that
change rank travel as packed device buffers through torch.distributed point-to-point calls (ncclSend / ncclRecv over
NVLink on the GPU box, gloo in the CPU tests).

The local energy of this harness is the kinetic energy -1/2 sum(L + G.G) (the Hamiltonian proper -- Coulomb, pseudopotentials
-- is out of scope); everything downstream of "an energy per walker" is the reference's arithmetic.

`engine` is duck-typed (qmcpack_b200.api.Crowd, or the oracle's VMC object in the parity tests):
    nw, capacity, dmc_sweep(), local_energies() -> [nw], rr() -> ([nw], [nw]), copy_walker(src, dst), set_num_walkers(n)
    and for multi-rank runs walker_bytes, pack_walker(iw, ptr), unpack_walker(iw, ptr).
"""
import math

import numpy as np

from . import sharding


class BranchEngine:
    """The subset of SFNBranch the DMC step needs (warm-up stage rules for E_trial; classic branch cut-off)."""

    def __init__(self, tau, target_walkers, e_trial=0.0, sigma2=10.0, target_sigma=10.0, feedback=1.0):
        self.tau = float(tau)
        self.tau_eff = float(tau)
        self.target = int(target_walkers)
        self.e_trial = self.e_ref = self.e_now = float(e_trial)
        self.feedback = float(feedback)
        self.r2_acc = self.r2_prop = 0.0
        self.set_branch_cutoff(sigma2, target_sigma, 50.0)

    def set_branch_cutoff(self, variance, target_sigma, max_sigma):
        """SFNBranch.cpp:290-320 ("classic"): min(max(variance * targetSigma, maxSigma), 2.5 / tau); max = 1.5 cut-off"""
        self.cutoff = min(max(variance * target_sigma, max_sigma), 2.5 / self.tau)
        self.bmax = 1.5 * self.cutoff
        self.bfilter = 1.0 / (self.bmax - self.cutoff)

    def branch_weight(self, enew, eold):
        """SFNBranch.h:199-208"""
        enew, eold = np.asarray(enew, np.float64), np.asarray(eold, np.float64)
        taueff = np.full_like(enew, 0.5 * self.tau_eff)
        x = np.maximum(self.e_ref - enew, self.e_ref - eold)
        taueff = np.where(x > self.bmax, 0.0, np.where(x > self.cutoff, taueff * (1.0 - (x - self.cutoff) * self.bfilter), taueff))
        return np.exp(taueff * (2.0 * self.e_trial - enew - eold))

    def update_after_pop_control(self, ens):
        """warm-up branch of SFNBranch::updateParamAfterPopControl (SFNBranch.cpp:133-199)"""
        self.e_now = ens["energy"]
        self.r2_acc += ens["r2_accepted"]
        self.r2_prop += ens["r2_proposed"]
        if self.r2_prop > 0:
            self.tau_eff = self.tau * self.r2_acc / self.r2_prop
        self.e_ref = self.e_now
        self.e_trial = self.e_now + (math.log(self.target) - math.log(ens["weight"])) / self.tau


def multiplicities(weights, uniforms):
    """WalkerControl.cpp:186: Multiplicity = static_cast<int>(Weight + rng())"""
    return np.floor(np.asarray(weights, np.float64) + np.asarray(uniforms, np.float64)).astype(np.int64)


def compute_cur_data(weights, energies, mult, rr_acc, rr_prop, rank=0, world=1, dist=None, device=None):
    """WalkerControl::computeCurData + the all-reduce; returns (ensemble dict, copies per rank)"""
    w, e = np.asarray(weights, np.float64), np.asarray(energies, np.float64)
    head = [float((w * e).sum()), float((w * e * e).sum()), float(len(w)), float(w.sum()), float(np.sum(rr_acc)),
            float(np.sum(rr_prop)), float((np.asarray(mult) > 0).sum())]
    per_rank = [0.0] * world
    per_rank[rank] = float(np.sum(mult))
    tot = sharding.reduce_block_estimator(head + per_rank, dist, device=device)
    esum, e2sum, nwalk, wsum, r2a, r2p, good = tot[:7]
    ens = dict(energy=esum / wsum, variance=e2sum / wsum - (esum / wsum) ** 2, weight=wsum, num_samples=nwalk,
               r2_accepted=r2a, r2_proposed=r2p, living_fraction=good / nwalk)
    return ens, [int(round(x)) for x in tot[7:]]


def swap_jobs(rank, num_per_rank, mult):
    """The send/receive schedule of WalkerControl::swapWalkersSimple (WalkerControl.cpp:330-415) for this rank.
    `mult` = multiplicities of this rank's walkers (dead ones are 0); returns (sends, mult_after) with
      sends: [(walker_index, target_rank, extra_copies)]
    (a receiver learns the number of folded copies from the message header: send_value(nsentcopy) in the reference)
    A sender always gives away its good walker with the most copies (ties: highest index, as std::sort on
    (copies, index) pairs leaves it last) and folds further copies bound for the same receiver into one message."""
    _, minus, plus = sharding.determine_new_walker_population(num_per_rank)
    mult = list(int(m) for m in mult)
    pairs = sorted((m, i) for i, m in enumerate(mult) if m > 0)
    sends = []
    nswap = len(plus)
    ic = 0
    while ic < nswap:
        nsent = 0
        if plus[ic] == rank:
            for idn in range(ic + 1, nswap):
                if plus[ic] == plus[idn] and minus[ic] == minus[idn] and pairs[-1][0] > 1:
                    pairs[-1] = (pairs[-1][0] - 1, pairs[-1][1])
                    nsent += 1
                else:
                    break
            sends.append((pairs[-1][1], minus[ic], nsent))
            if pairs[-1][0] > 1:
                pairs[-1] = (pairs[-1][0] - 1, pairs[-1][1])
                pairs.sort()
            else:
                mult[pairs[-1][1]] = 0
                pairs.pop()
        ic += 1 + nsent
    for m, i in pairs:
        mult[i] = m
    return sends, mult


class DMC:
    """One rank's DMC population on one crowd (the crowd's capacity bounds the local population)."""

    def __init__(self, engine, tau, target_walkers, branch_rng, dist=None, device=None):
        self.eng = engine
        self.branch = BranchEngine(tau, target_walkers)
        self.rng = branch_rng  # callable returning one uniform: WalkerControl's own generator
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        n = engine.nw
        self.weights = np.ones(n)
        self.ages = np.zeros(n, np.int64)
        self.energies = np.asarray(engine.local_energies(), np.float64).copy()
        self.branch.e_trial = self.branch.e_ref = float(self.energies.mean())
        self.rr_acc = np.zeros(n)
        self.rr_prop = np.zeros(n)
        self.history = []

    # ---- DMCBatched::advanceWalkers
    def advance(self):
        e_old = self.energies
        self.eng.dmc_sweep()
        e_new = np.asarray(self.eng.local_energies(), np.float64)
        self.rr_acc, self.rr_prop = (np.asarray(x, np.float64) for x in self.eng.rr())
        self.weights = self.weights * self.branch.branch_weight(e_new, e_old)  # DMCBatched.cpp:284-285
        self.ages = np.where(self.rr_prop > 0, 0, self.ages + 1)               # :286-289
        self.energies = e_new.copy()

    # ---- WalkerControl::branch, dynamic population
    def branch_step(self, do_not_branch=False):
        eng, n = self.eng, self.eng.nw
        if do_not_branch:
            mult = np.ones(n, np.int64)
        else:
            mult = multiplicities(self.weights, [self.rng() for _ in range(n)])
        ens, num_per_rank = compute_cur_data(self.weights, self.energies, mult, self.rr_acc, self.rr_prop, self.rank,
                                             self.world, self.dist, self.device)
        incoming = []
        if self.world > 1:
            mult, incoming = self._swap(mult, num_per_rank)
        mult = list(int(m) for m in mult)
        # killDeadWalkersOnRank: survivors keep their order and close the gaps (vector erase); received walkers were
        # appended by spawnWalker during the swap; fissionHighMultiplicityWalkers then walks the whole list once and
        # appends the copies parent by parent (survivors' copies first, then the copies of received walkers)
        survivors = [i for i in range(len(mult)) if mult[i] > 0]
        new_w, new_e, new_age, new_mult = [], [], [], []
        for dst, src in enumerate(survivors):
            eng.copy_walker(src, dst)
            new_w.append(self.weights[src])
            new_e.append(self.energies[src])
            new_age.append(self.ages[src])
            new_mult.append(mult[src])
        nlive = len(survivors)
        for buf, w_in, e_in, age_in, extra in incoming:
            if nlive >= eng.capacity:
                raise RuntimeError("DMC population exceeds the crowd's capacity on rank %d" % self.rank)
            eng.unpack_walker(nlive, buf.data_ptr())
            new_w.append(w_in)
            new_e.append(e_in)
            new_age.append(age_in)
            new_mult.append(extra + 1)
            nlive += 1
        for parent in range(len(new_mult)):
            for _ in range(new_mult[parent] - 1):
                if nlive >= eng.capacity:
                    raise RuntimeError("DMC population exceeds the crowd's capacity on rank %d" % self.rank)
                eng.copy_walker(parent, nlive)
                new_w.append(new_w[parent])
                new_e.append(new_e[parent])
                new_age.append(new_age[parent])
                nlive += 1
        if nlive == 0:
            raise RuntimeError("DMC population died out on rank %d" % self.rank)
        eng.set_num_walkers(nlive)
        self.energies = np.asarray(new_e, np.float64)
        self.ages = np.asarray(new_age, np.int64)
        # WalkerControl.cpp:226-233: Weight and Multiplicity go back to 1 only `if (!do_not_branch)`; on iteration 0 the
        # branch weights of the first step carry into the second
        self.weights = np.asarray(new_w, np.float64) if do_not_branch else np.ones(nlive)
        self.rr_acc = np.zeros(nlive)
        self.rr_prop = np.zeros(nlive)
        self.branch.update_after_pop_control(ens)
        self.history.append(dict(ens, population=sum(num_per_rank), local=nlive, e_trial=self.branch.e_trial))
        return ens

    def _swap(self, mult, num_per_rank):
        """swapWalkersSimple: packed device buffers over point-to-point messages; header = (extra copies, weight, energy,
        age) -- Weight, Age and the Properties travel inside Walker::DataSet in the reference"""
        import torch
        dist, eng = self.dist, self.eng
        _, minus, plus = sharding.determine_new_walker_population(num_per_rank)
        sends, mult_after = swap_jobs(self.rank, num_per_rank, mult)
        dev = self.device if self.device is not None else "cpu"
        nbytes = eng.walker_bytes
        incoming = []
        send_iter = iter(sends)
        ic = 0
        while ic < len(plus):
            nsent = 0
            if plus[ic] == self.rank:
                widx, target, nsent = next(send_iter)
                head = torch.tensor([float(nsent), float(self.weights[widx]), float(self.energies[widx]),
                                     float(self.ages[widx])], dtype=torch.float64, device=dev)
                buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                eng.pack_walker(widx, buf.data_ptr())
                dist.send(head, dst=target)
                dist.send(buf, dst=target)
            elif minus[ic] == self.rank:
                head = torch.empty(4, dtype=torch.float64, device=dev)
                buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                dist.recv(head, src=plus[ic])
                dist.recv(buf, src=plus[ic])
                h = head.cpu().numpy()
                nsent = int(h[0])
                incoming.append((buf, float(h[1]), float(h[2]), int(h[3]), nsent))
            else:
                # a third rank must skip the folded entries too: the fold count is (pairs with the same endpoints that
                # follow) limited by the sender's copies; it is broadcast implicitly by advancing one entry at a time,
                # which is safe because folded entries name the same (sender, receiver) pair this rank is not part of
                nsent = 0
            ic += 1 + nsent
        return mult_after, incoming
