// qmcpack_b200/csrc/dmc_host.cpp -- DMC layer above the device-resident sweep (include/qmcb_driver.h).  Plain C++17.
//
// Restates (paths under /root/reference/src/QMCDrivers):
//   DMC/DMCBatched.cpp:264-292     branch weights and ages after the move loop
//   SFNBranch.h:199-208            branchWeight;  SFNBranch.cpp:133-215 updateParamAfterPopControl (warm-up and main stage);
//   SFNBranch.cpp:290-320          setBranchCutoff ("classic")
//   DMC/WalkerControl.cpp:151-240  branch;  :243-281 computeCurData;  :284-313 determineNewWalkerPopulation;
//   DMC/WalkerControl.cpp:312-500  swapWalkersSimple;  :539-553 killDeadWalkersOnRank
//   MCPopulation.cpp:34-66         fissionHighMultiplicityWalkers
//   Utilities/FairDivide.h:129-145 FairDivideLow
#include "../../include/qmcb_driver.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace
{
thread_local std::string g_dmc_err;

// ---- engine over a qmcb_crowd (the harness' local energy is the kinetic energy, qmcb_twf_mw_evaluate_gl)
struct CrowdEngine
{
  qmcb_crowd* c;
};
int ce_num(void* x) { return qmcb_crowd_num_walkers(static_cast<CrowdEngine*>(x)->c); }
int ce_cap(void* x) { return qmcb_crowd_capacity(static_cast<CrowdEngine*>(x)->c); }
int ce_sweep(void* x) { return qmcb_vmc_sweep(static_cast<CrowdEngine*>(x)->c, 1, nullptr); }
int ce_energies(void* x, double* e) { return qmcb_twf_mw_evaluate_gl(static_cast<CrowdEngine*>(x)->c, nullptr, nullptr, nullptr, e); }
int ce_rr(void* x, double* a, double* p) { return qmcb_dmc_get_rr(static_cast<CrowdEngine*>(x)->c, a, p); }
int ce_copy(void* x, int s, int d) { return qmcb_crowd_copy_walker(static_cast<CrowdEngine*>(x)->c, s, d); }
int ce_setn(void* x, int n) { return qmcb_crowd_set_num_walkers(static_cast<CrowdEngine*>(x)->c, n); }
int ce_pack(void* x, int iw, void* b) { return qmcb_crowd_pack_walker(static_cast<CrowdEngine*>(x)->c, iw, b); }
int ce_unpack(void* x, int iw, const void* b) { return qmcb_crowd_unpack_walker(static_cast<CrowdEngine*>(x)->c, iw, b); }

// accumulator_set<T>: result() = mean of the pushed values
struct Accumulator
{
  double sum = 0;
  long long n = 0;
  void operator()(double v)
  {
    sum += v;
    ++n;
  }
  double result() const { return n ? sum / (double)n : 0.0; }
  double mean() const { return result(); }
  long long count() const { return n; }
};

// curData layout (WalkerControl.h: ENERGY_INDEX ...)
enum
{
  ENERGY_INDEX = 0,
  ENERGY_SQ_INDEX,
  WALKERSIZE_INDEX,
  WEIGHT_INDEX,
  R2ACCEPTED_INDEX,
  R2PROPOSED_INDEX,
  FNSIZE_INDEX,
  SENTWALKERS_INDEX,
  LE_MAX
};

struct Walker
{
  double weight = 1.0, energy = 0.0, rr_acc = 0.0, rr_prop = 0.0;
  long long age = 0;
  int mult = 1;
  int slot = 0; // where the walker's state lives in the engine right now
};
} // namespace

struct qmcb_dmc
{
  qmcb_dmc_engine eng{};
  CrowdEngine crowd_engine{};
  qmcb_comm comm{};
  bool has_comm = false;
  qmcb_dmc_params prm{};
  // ---- SFNBranch
  double tau = 0, tau_eff = 0, e_trial = 0, e_ref = 0, e_now = 0, sigma2 = 10, sigma_bound = 10, feedback = 1;
  double cutoff = 0, bmax = 0, bfilter = 0;
  int target_walkers = 0, warmup_todo = 0, etrial_update_todo = 1;
  bool dmc_stage = false;
  Accumulator r2_acc, r2_prop, energy_hist, variance_hist;
  // ---- WalkerControl
  std::mt19937 rng;
  long long saved_num_walkers_sent = 0;
  std::vector<int> num_per_rank;
  // ---- population (live walkers in engine order)
  std::vector<Walker> w;

  double uniform()
  {
    // Utilities/StdRandom.h:34-48
    return static_cast<double>(rng() - rng.min()) / (static_cast<double>(rng.max() - rng.min()) + 1);
  }
  void set_branch_cutoff(double variance, double target_sigma, double max_sigma)
  {
    cutoff  = std::min(std::max(variance * target_sigma, max_sigma), 2.5 / tau);
    bmax    = cutoff * 1.5;
    bfilter = 1.0 / (bmax - cutoff);
  }
  double branch_weight(double enew, double eold) const
  {
    double taueff_ = tau_eff * 0.5;
    const double x = std::max(e_ref - enew, e_ref - eold);
    if (x > bmax)
      taueff_ = 0.0;
    else if (x > cutoff)
      taueff_ *= (1.0 - (x - cutoff) * bfilter);
    return std::exp(taueff_ * (e_trial * 2.0 - enew - eold));
  }

  void init()
  {
    const int n = eng.num_walkers(eng.ctx);
    if (n < 1)
      throw std::runtime_error("qmcb_dmc: the engine has no live walkers");
    w.assign(n, Walker());
    std::vector<double> e(n);
    if (eng.local_energies(eng.ctx, e.data()))
      throw std::runtime_error("qmcb_dmc: local_energies failed");
    for (int i = 0; i < n; ++i)
    {
      w[i].energy = e[i];
      w[i].slot   = i;
    }
    tau = tau_eff = prm.tau;
    sigma2        = prm.sigma2 > 0 ? prm.sigma2 : 10.0;
    sigma_bound   = prm.sigma_bound > 0 ? prm.sigma_bound : 10.0;
    feedback      = prm.feedback > 0 ? prm.feedback : 1.0;
    warmup_todo   = prm.warmup_steps;
    dmc_stage     = prm.warmup_steps == 0; // BranchMode.set(B_DMCSTAGE, iParam[B_WARMUPSTEPS] == 0)
    etrial_update_todo = std::max(1, prm.energy_update_interval);
    rng.seed(prm.branch_seed);
    // SFNBranch::initParam: E_trial = E_ref = <E> over the GLOBAL initial population
    double acc[2] = {0.0, (double)n};
    for (int i = 0; i < n; ++i)
      acc[0] += e[i];
    if (has_comm && comm.size > 1 && comm.allreduce_sum(comm.ctx, acc, 2))
      throw std::runtime_error("qmcb_dmc: allreduce failed");
    e_trial = e_ref = e_now = acc[0] / acc[1];
    target_walkers          = prm.target_walkers > 0 ? prm.target_walkers : (int)std::llround(acc[1]);
    set_branch_cutoff(sigma2, sigma_bound, 50.0);
    num_per_rank.assign(has_comm ? comm.size : 1, 0);
  }

  // ---- DMCBatched::advanceWalkers after the move loop (DMCBatched.cpp:264-292)
  void advance()
  {
    const int n = (int)w.size();
    if (eng.sweep(eng.ctx))
      throw std::runtime_error(std::string("qmcb_dmc: sweep failed: ") + qmcb_last_error());
    std::vector<double> e(n), ra(n), rp(n);
    if (eng.local_energies(eng.ctx, e.data()) || eng.get_rr(eng.ctx, ra.data(), rp.data()))
      throw std::runtime_error(std::string("qmcb_dmc: reading energies / rr failed: ") + qmcb_last_error());
    for (int i = 0; i < n; ++i)
    {
      w[i].weight *= branch_weight(e[i], w[i].energy);
      w[i].age = rp[i] > 0 ? 0 : w[i].age + 1;
      w[i].energy  = e[i];
      w[i].rr_acc  = ra[i];
      w[i].rr_prop = rp[i];
    }
  }

  static void fair_divide_low(int ntot, int npart, std::vector<int>& off)
  {
    off.assign(npart + 1, 0);
    const int bat = ntot / npart, residue = npart - ntot % npart;
    for (int i = 0; i < npart; ++i)
      off[i + 1] = off[i] + (i < residue ? bat : bat + 1);
  }

  // WalkerControl::killDeadWalkersOnRank: drop the walkers with Multiplicity 0; survivors keep their order (vector erase)
  // and the engine's slots follow at once: walker i lives in slot i again afterwards.  Slots only ever move DOWN here
  // (a survivor's slot is >= its new index and everything below is already in place), so no spare slot is needed.
  void kill_dead()
  {
    std::vector<Walker> live;
    live.reserve(w.size());
    for (auto& x : w)
      if (x.mult != 0)
      {
        const int dst = (int)live.size();
        if (x.slot != dst)
        {
          if (eng.copy_walker(eng.ctx, x.slot, dst))
            throw std::runtime_error(std::string("qmcb_dmc: copy_walker failed: ") + qmcb_last_error());
          x.slot = dst;
        }
        live.push_back(x);
      }
    w.swap(live);
  }
  // MCPopulation::spawnWalker: the next slot behind the population (`pending` walkers are already parked there)
  int take_slot(size_t pending = 0)
  {
    const size_t s = w.size() + pending;
    if ((int)s >= eng.capacity(eng.ctx))
      throw std::runtime_error("qmcb_dmc: the population exceeds the crowd's capacity (" +
                               std::to_string(eng.capacity(eng.ctx)) + " walkers on rank " +
                               std::to_string(has_comm ? comm.rank : 0) + "); create the crowd with more head-room");
    return (int)s;
  }

  // ---- WalkerControl::branch(iter, pop, do_not_branch), dynamic population
  void branch(int /*iter*/, bool do_not_branch, qmcb_dmc_ensemble* out)
  {
    const int nranks = has_comm ? comm.size : 1, rank = has_comm ? comm.rank : 0;
    for (auto& x : w)
      x.mult = do_not_branch ? 1 : static_cast<int>(x.weight + uniform());
    // computeCurData
    std::vector<double> cur(LE_MAX + nranks, 0.0);
    {
      double esum = 0, e2sum = 0, wsum = 0, r2a = 0, r2p = 0;
      int good = 0, copies = 0;
      for (auto& x : w)
      {
        good += x.mult > 0 ? 1 : 0;
        copies += x.mult;
        r2a += x.rr_acc;
        r2p += x.rr_prop;
        esum += x.weight * x.energy;
        e2sum += x.weight * x.energy * x.energy;
        wsum += x.weight;
      }
      cur[ENERGY_INDEX]      = esum;
      cur[ENERGY_SQ_INDEX]   = e2sum;
      cur[WALKERSIZE_INDEX]  = (double)w.size();
      cur[WEIGHT_INDEX]      = wsum;
      cur[R2ACCEPTED_INDEX]  = r2a;
      cur[R2PROPOSED_INDEX]  = r2p;
      cur[FNSIZE_INDEX]      = good;
      cur[SENTWALKERS_INDEX] = (double)saved_num_walkers_sent;
      cur[LE_MAX + rank]     = copies;
      if (nranks > 1 && comm.allreduce_sum(comm.ctx, cur.data(), (int)cur.size()))
        throw std::runtime_error("qmcb_dmc: allreduce failed");
    }
    for (int i = 0; i < nranks; ++i)
      num_per_rank[i] = static_cast<int>(cur[LE_MAX + i]);
    // writeDMCdat: ensemble properties of this generation
    qmcb_dmc_ensemble ens{};
    const double wgt_inv = 1.0 / cur[WEIGHT_INDEX];
    ens.energy           = cur[ENERGY_INDEX] * wgt_inv;
    ens.weight           = cur[WEIGHT_INDEX];
    ens.variance         = cur[ENERGY_SQ_INDEX] * wgt_inv - ens.energy * ens.energy;
    ens.num_samples      = cur[WALKERSIZE_INDEX];
    ens.r2_accepted      = cur[R2ACCEPTED_INDEX];
    ens.r2_proposed      = cur[R2PROPOSED_INDEX];
    ens.living_fraction  = cur[FNSIZE_INDEX] / cur[WALKERSIZE_INDEX];

    kill_dead();
    if (nranks > 1)
      swap_walkers_simple(ens);
    kill_dead();
    // MCPopulation::fissionHighMultiplicityWalkers: copies are appended parent by parent
    {
      const size_t nparents = w.size();
      for (size_t i = 0; i < nparents; ++i)
      {
        int num_copies = w[i].mult;
        while (num_copies > 1)
        {
          Walker c = w[i];
          c.slot   = take_slot();
          c.mult   = 1;
          if (eng.copy_walker(eng.ctx, w[i].slot, c.slot))
            throw std::runtime_error(std::string("qmcb_dmc: copy_walker failed: ") + qmcb_last_error());
          w.push_back(c);
          w[i].mult -= 1;
          --num_copies;
        }
      }
    }
    if (w.empty())
      throw std::runtime_error("qmcb_dmc: the population died out on rank " + std::to_string(rank));
    if (eng.set_num_walkers(eng.ctx, (int)w.size()))
      throw std::runtime_error(std::string("qmcb_dmc: set_num_walkers failed: ") + qmcb_last_error());
    if (!do_not_branch)
      for (auto& x : w)
      {
        x.weight = 1.0;
        x.mult   = 1;
      }
    ens.population = std::accumulate(num_per_rank.begin(), num_per_rank.end(), 0);
    ens.local      = (int)w.size();
    update_param_after_pop_control(ens);
    ens.e_trial       = e_trial;
    ens.e_ref         = e_ref;
    ens.tau_eff       = tau_eff;
    ens.branch_cutoff = cutoff;
    if (out)
      *out = ens;
  }

  // ---- WalkerControl::swapWalkersSimple: every rank derives the same schedule from num_per_rank
  void swap_walkers_simple(qmcb_dmc_ensemble& ens)
  {
    const int nranks = comm.size, rank = comm.rank;
    std::vector<int> fair, minus, plus;
    const int pop = std::accumulate(num_per_rank.begin(), num_per_rank.end(), 0);
    fair_divide_low(pop, nranks, fair);
    for (int ip = 0; ip < nranks; ++ip)
    {
      const int dn = num_per_rank[ip] - (fair[ip + 1] - fair[ip]);
      if (dn > 0)
        plus.insert(plus.end(), dn, ip);
      else if (dn < 0)
        minus.insert(minus.end(), -dn, ip);
    }
    if (plus.size() != minus.size())
      throw std::runtime_error("qmcb_dmc: walker send/recv pattern does not match");
    const int nswap = (int)plus.size();
    std::vector<std::pair<int, int>> ncopy_pairs; // (multiplicity, index), sorted: the back has most copies
    for (int iw = 0; iw < (int)w.size(); ++iw)
      ncopy_pairs.push_back(std::make_pair(w[iw].mult, iw));
    std::sort(ncopy_pairs.begin(), ncopy_pairs.end());
    long long nsend = 0;
    std::vector<Walker> received;
    for (int ic = 0; ic < nswap; ++ic)
    {
      int nsentcopy = 0;
      if (plus[ic] == rank)
      {
        // always send the last good walker with most copies; fold further copies for the same receiver into one message
        for (int id = ic + 1; id < nswap; ++id)
          if (plus[ic] == plus[id] && minus[ic] == minus[id] && ncopy_pairs.back().first > 1)
          {
            ncopy_pairs.back().first--;
            nsentcopy++;
          }
          else
            break;
        const Walker& src = w[ncopy_pairs.back().second];
        if (eng.pack_walker(eng.ctx, src.slot, comm.send_buf))
          throw std::runtime_error(std::string("qmcb_dmc: pack_walker failed: ") + qmcb_last_error());
        const double header[4] = {(double)nsentcopy, src.weight, src.energy, (double)src.age};
        if (comm.send(comm.ctx, minus[ic], header))
          throw std::runtime_error("qmcb_dmc: send failed");
        ++nsend;
        if (ncopy_pairs.back().first > 1)
        {
          ncopy_pairs.back().first--;
          std::sort(ncopy_pairs.begin(), ncopy_pairs.end());
        }
        else
        {
          w[ncopy_pairs.back().second].mult = 0;
          ncopy_pairs.pop_back();
        }
      }
      if (minus[ic] == rank)
      {
        double header[4];
        if (comm.recv(comm.ctx, plus[ic], header))
          throw std::runtime_error("qmcb_dmc: recv failed");
        nsentcopy = (int)header[0];
        if (ic + nsentcopy >= nswap || plus[ic] != plus[ic + nsentcopy] || minus[ic] != minus[ic + nsentcopy])
          throw std::runtime_error("qmcb_dmc: send/recv pair checking failed");
        Walker nw;
        nw.slot   = take_slot(received.size());
        nw.weight = header[1];
        nw.energy = header[2];
        nw.age    = (long long)header[3];
        nw.mult   = nsentcopy + 1;
        if (eng.unpack_walker(eng.ctx, nw.slot, comm.recv_buf))
          throw std::runtime_error(std::string("qmcb_dmc: unpack_walker failed: ") + qmcb_last_error());
        received.push_back(nw);
        ens.walkers_received += 1;
      }
      ic += nsentcopy;
    }
    saved_num_walkers_sent = nsend;
    ens.walkers_sent       = nsend;
    // rebuild Multiplicity
    for (auto& pr : ncopy_pairs)
      w[pr.second].mult = pr.first;
    for (auto& r : received)
      w.push_back(r);
  }

  // ---- SFNBranch::updateParamAfterPopControl (SFNBranch.cpp:133-215), B_POPCONTROL on, B_KILLNODES off
  void update_param_after_pop_control(const qmcb_dmc_ensemble& ens)
  {
    const double logN       = std::log((double)target_walkers);
    const double pop_weight = ens.weight;
    e_now                   = ens.energy;
    r2_acc(ens.r2_accepted);
    r2_prop(ens.r2_proposed);
    if (prm.use_tau_eff && r2_prop.result() > 0)
      tau_eff = tau * r2_acc.result() / r2_prop.result();
    if (dmc_stage)
    {
      energy_hist(e_now);
      variance_hist(ens.variance);
      e_ref = energy_hist.mean();
      --etrial_update_todo;
      if (etrial_update_todo == 0)
      {
        e_trial            = e_ref + feedback * (logN - std::log(pop_weight));
        etrial_update_todo = std::max(1, prm.energy_update_interval);
      }
    }
    else
    {
      e_ref   = e_now;
      e_trial = e_now + (logN - std::log(pop_weight)) / tau;
      --warmup_todo;
      if (warmup_todo == 0)
      {
        sigma2 = ens.variance;
        set_branch_cutoff(sigma2, sigma_bound, 10.0);
        dmc_stage = true;
      }
    }
  }
};

namespace
{
template<typename F>
int guarded_dmc(F&& f)
{
  try
  {
    f();
    return 0;
  }
  catch (const std::exception& e)
  {
    g_dmc_err = e.what();
    return 1;
  }
}
} // namespace

extern "C" {
const char* qmcb_dmc_last_error(void) { return g_dmc_err.c_str(); }

int qmcb_dmc_create_with_engine(qmcb_dmc** d, const qmcb_dmc_engine* engine, const qmcb_dmc_params* p, const qmcb_comm* comm)
{
  return guarded_dmc([&] {
    if (!d || !engine || !p)
      throw std::runtime_error("qmcb_dmc_create: null argument");
    if (!(p->tau > 0))
      throw std::runtime_error("qmcb_dmc_create: tau must be positive");
    auto* x = new qmcb_dmc();
    x->eng  = *engine;
    x->prm  = *p;
    if (comm && comm->size > 1)
    {
      if (!comm->allreduce_sum || !comm->send || !comm->recv || !comm->send_buf || !comm->recv_buf)
      {
        delete x;
        throw std::runtime_error("qmcb_dmc_create: incomplete communicator");
      }
      x->comm     = *comm;
      x->has_comm = true;
    }
    try
    {
      x->init();
    }
    catch (...)
    {
      delete x;
      throw;
    }
    *d = x;
  });
}

int qmcb_dmc_create(qmcb_dmc** d, qmcb_crowd* crowd, const qmcb_dmc_params* p, const qmcb_comm* comm)
{
  return guarded_dmc([&] {
    if (!crowd)
      throw std::runtime_error("qmcb_dmc_create: null crowd");
    qmcb_dmc_engine e{};
    e.num_walkers     = ce_num;
    e.capacity        = ce_cap;
    e.sweep           = ce_sweep;
    e.local_energies  = ce_energies;
    e.get_rr          = ce_rr;
    e.copy_walker     = ce_copy;
    e.set_num_walkers = ce_setn;
    e.pack_walker     = ce_pack;
    e.unpack_walker   = ce_unpack;
    qmcb_dmc* x       = nullptr;
    // the engine context must outlive the call: it is a member of the object, patched in after construction
    CrowdEngine tmp{crowd};
    e.ctx = &tmp;
    if (qmcb_dmc_create_with_engine(&x, &e, p, comm))
      throw std::runtime_error(g_dmc_err);
    x->crowd_engine = tmp;
    x->eng.ctx      = &x->crowd_engine;
    *d              = x;
  });
}

int qmcb_dmc_destroy(qmcb_dmc* d)
{
  delete d;
  return 0;
}

int qmcb_dmc_advance(qmcb_dmc* d)
{
  return guarded_dmc([&] {
    if (!d)
      throw std::runtime_error("null handle");
    d->advance();
  });
}

int qmcb_dmc_branch(qmcb_dmc* d, int iter, int do_not_branch, qmcb_dmc_ensemble* out)
{
  return guarded_dmc([&] {
    if (!d)
      throw std::runtime_error("null handle");
    d->branch(iter, do_not_branch != 0, out);
  });
}

int qmcb_dmc_step(qmcb_dmc* d, int iter, qmcb_dmc_ensemble* out)
{
  return guarded_dmc([&] {
    if (!d)
      throw std::runtime_error("null handle");
    d->advance();
    d->branch(iter, iter == 0, out); // DMCBatched.cpp:512
  });
}

int qmcb_dmc_get_walkers(qmcb_dmc* d, double* weights, double* energies, long long* ages, int max_n)
{
  if (!d)
    return -1;
  const int n = (int)d->w.size();
  for (int i = 0; i < n && i < max_n; ++i)
  {
    if (weights)
      weights[i] = d->w[i].weight;
    if (energies)
      energies[i] = d->w[i].energy;
    if (ages)
      ages[i] = d->w[i].age;
  }
  return n;
}

int qmcb_dmc_set_weights(qmcb_dmc* d, const double* weights, int n)
{
  return guarded_dmc([&] {
    if (!d || !weights || n != (int)d->w.size())
      throw std::runtime_error("qmcb_dmc_set_weights: size mismatch");
    for (int i = 0; i < n; ++i)
      d->w[i].weight = weights[i];
  });
}

double qmcb_dmc_branch_weight(qmcb_dmc* d, double enew, double eold) { return d ? d->branch_weight(enew, eold) : 0.0; }
}
