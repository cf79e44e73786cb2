// qmcpack_b200/csrc/vmc_host.cpp -- host driver above the C ABI (see include/qmcb_driver.h).  Plain C++17, no CUDA.
#include "../../include/qmcb_driver.h"
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <stdexcept>
#include <string>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>

namespace
{
thread_local std::string g_err;

// Utilities/StdRandom.h:34-48: (eng() - min) / (max - min + 1) in double
struct StdRandom
{
  std::mt19937 engine;
  explicit StdRandom(uint32_t seed) : engine(seed) {}
  double operator()()
  {
    return static_cast<double>(engine() - engine.min()) / (static_cast<double>(engine.max() - engine.min()) + 1);
  }
};

// RandomSeqGenerator.h:33-52
template<typename T>
void assignGaussRand(T* a, unsigned n, StdRandom& rng)
{
  const double slightly_less_than_one = 1.0 - std::numeric_limits<double>::epsilon();
  const int nm1                       = n - 1;
  double temp1, temp2;
  for (int i = 0; i < nm1; i += 2)
  {
    temp1    = std::sqrt(-2.0 * std::log(1.0 - slightly_less_than_one * rng()));
    temp2    = 2.0 * M_PI * rng();
    a[i]     = temp1 * std::cos(temp2);
    a[i + 1] = temp1 * std::sin(temp2);
  }
  if (n % 2 == 1)
  {
    temp1  = std::sqrt(-2.0 * std::log(1.0 - slightly_less_than_one * rng()));
    temp2  = 2.0 * M_PI * rng();
    a[nm1] = temp1 * std::cos(temp2);
  }
}

// DriftModifierUNR.cpp:20-31, a = 1
template<typename RT>
void getDrift(RT tau, const RT qf[3], RT drift[3])
{
  const RT a   = RT(1);
  const RT vsq = qf[0] * qf[0] + qf[1] * qf[1] + qf[2] * qf[2];
  const RT sc =
      vsq < std::numeric_limits<RT>::epsilon() ? tau : ((-1.0 + std::sqrt(1.0 + 2.0 * a * tau * vsq)) / (a * vsq));
  for (int d = 0; d < 3; ++d)
    drift[d] = qf[d] * sc;
}

void chk(int rc)
{
  if (rc != 0)
    throw std::runtime_error(qmcb_last_error());
}

struct CrowdCtx
{
  qmcb_crowd* crowd;
  int nw, w0;
  StdRandom rng;
  long long n_accept = 0, n_reject = 0;
  CrowdCtx(qmcb_crowd* c, int n, int off, uint32_t seed) : crowd(c), nw(n), w0(off), rng(seed) {}
};

// QMCB_HOST_PROFILE=1: wall-clock attribution of the host-driven loop (crowd 0 prints microseconds per move spent in
// each C-ABI call and in the host arithmetic between them) -- where a round trip goes when no timeline tool is at hand
struct PhaseClock
{
  bool on;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point t;
  explicit PhaseClock(bool enable) : on(enable)
  {
    if (on)
      t = std::chrono::steady_clock::now();
  }
  void lap(int phase)
  {
    if (!on)
      return;
    const auto n = std::chrono::steady_clock::now();
    acc[phase] += std::chrono::duration<double, std::micro>(n - t).count();
    t = n;
  }
};

template<typename RT>
void advance_crowd(CrowdCtx& cx, int N, double tau, bool use_drift, uint8_t* log, int nw_total)
{
  static const bool profile = [] {
    const char* e = std::getenv("QMCB_HOST_PROFILE");
    return e && std::atoi(e) != 0;
  }();
  PhaseClock pc(profile && cx.w0 == 0);
  const int nw = cx.nw;
  // TauParams.hpp:29-40, unit mass
  const RT tauovermass = RT(tau) * RT(1.0);
  const RT oneover2tau = 0.5 / (tauovermass);
  const RT sqrttau     = std::sqrt(tauovermass);
  std::vector<RT> walker_deltas(3 * (size_t)nw * N), deltas(3 * (size_t)nw), drifts(3 * (size_t)nw), log_gf(nw, RT(0)),
      log_gb(nw, RT(0)), prob(nw);
  // complex orbitals: ratios [nw][2] and grads [nw][3][2] arrive as interleaved (re, im) doubles; the drift takes the
  // real part of the gradient (DriftModifierUNR.cpp:20-23) and prob = std::norm(ratio) (VMCBatched.cpp:152)
  const int cs = qmcb_crowd_is_complex(cx.crowd) ? 2 : 1;
  std::vector<double> grads(3 * (size_t)nw * cs), displ(3 * (size_t)nw), ratios((size_t)nw * cs);
  std::vector<uint8_t> accepted(nw);
  std::vector<double> uni(nw);
  assignGaussRand(walker_deltas.data(), (unsigned)walker_deltas.size(), cx.rng);
  for (int iat = 0; iat < N; ++iat)
  {
    // Everything that does not depend on the device's answers runs HERE, while the device still applies the previous
    // accept and prepares this electron's row: the scaled Gaussians, log_gf (a function of the Gaussians only) and the
    // uniforms of the accept tests, drawn ahead from a COPY of the crowd's engine -- the reference draws walker i's
    // uniform only when prob[i] >= eps (VMCBatched.cpp:156-158), so the copy is committed only when every walker of the
    // move turns out to draw; otherwise the tests are replayed from the untouched engine, walker by walker.
    for (int i = 0; i < nw; ++i)
      for (int d = 0; d < 3; ++d)
        deltas[3 * i + d] = walker_deltas[3 * ((size_t)iat * nw + i) + d] * sqrttau;
    if (use_drift)
      for (int i = 0; i < nw; ++i)
      {
        const RT* dl = &deltas[3 * i];
        log_gf[i]    = -oneover2tau * (dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
      }
    StdRandom ahead = cx.rng;
    for (int i = 0; i < nw; ++i)
      uni[i] = ahead();
    pc.lap(0);
    if (use_drift)
    {
      chk(qmcb_twf_mw_eval_grad(cx.crowd, iat, grads.data()));
      pc.lap(1);
      for (int i = 0; i < nw; ++i)
      {
        const RT g[3] = {(RT)grads[(3 * i) * cs], (RT)grads[(3 * i + 1) * cs], (RT)grads[(3 * i + 2) * cs]};
        RT dr[3];
        getDrift<RT>(tauovermass, g, dr);
        for (int d = 0; d < 3; ++d)
          drifts[3 * i + d] = dr[d] + deltas[3 * i + d];
      }
    }
    else
      drifts = deltas;
    for (size_t i = 0; i < displ.size(); ++i)
      displ[i] = drifts[i];
    pc.lap(0);
    chk(qmcb_ps_mw_make_move(cx.crowd, iat, displ.data()));
    pc.lap(2);
    chk(qmcb_twf_mw_calc_ratio_grad(cx.crowd, iat, ratios.data(), grads.data()));
    pc.lap(3);
    if (use_drift)
      for (int i = 0; i < nw; ++i)
      {
        const RT g[3] = {(RT)grads[(3 * i) * cs], (RT)grads[(3 * i + 1) * cs], (RT)grads[(3 * i + 2) * cs]};
        RT dr[3];
        getDrift<RT>(tauovermass, g, dr);
        for (int d = 0; d < 3; ++d)
          dr[d] += drifts[3 * i + d];
        log_gb[i] = -oneover2tau * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
      }
    for (int i = 0; i < nw; ++i)
      prob[i] = cs == 2 ? (RT)(ratios[2 * i] * ratios[2 * i] + ratios[2 * i + 1] * ratios[2 * i + 1]) : (RT)(ratios[i] * ratios[i]);
    bool all_draw = true;
    for (int i = 0; i < nw; ++i)
      all_draw = all_draw && prob[i] >= std::numeric_limits<RT>::epsilon();
    if (all_draw)
      cx.rng = ahead; // every walker consumed exactly the uniform drawn ahead for it
    for (int i = 0; i < nw; ++i)
    {
      // the uniform is drawn only when the move is valid and prob >= eps (VMCBatched.cpp:156-158)
      if (prob[i] >= std::numeric_limits<RT>::epsilon() &&
          (all_draw ? uni[i] : cx.rng()) < prob[i] * std::exp(log_gb[i] - log_gf[i]))
      {
        accepted[i] = 1;
        cx.n_accept++;
      }
      else
      {
        accepted[i] = 0;
        cx.n_reject++;
      }
      if (log)
        log[(size_t)iat * nw_total + cx.w0 + i] = accepted[i];
    }
    pc.lap(0);
    chk(qmcb_twf_mw_accept_reject(cx.crowd, iat, accepted.data(), 1));
    pc.lap(4);
  }
  if (pc.on)
    std::fprintf(stderr,
                 "[host profile] crowd 0 (%d walkers), us per move: host arithmetic %.1f | eval_grad %.1f | make_move %.1f | "
                 "calc_ratio_grad %.1f | accept_reject %.1f\n",
                 nw, pc.acc[0] / N, pc.acc[1] / N, pc.acc[2] / N, pc.acc[3] / N, pc.acc[4] / N);
  chk(qmcb_twf_mw_complete_updates(cx.crowd));
  chk(qmcb_crowd_sync(cx.crowd));
}
} // namespace

struct qmcb_host_vmc
{
  std::vector<CrowdCtx> crowds;
  int N = 0, precision = 0, nw_total = 0;
  double tau    = 0.3;
  bool use_drift = true;
};

extern "C"
{
const char* qmcb_host_vmc_last_error(void) { return g_err.c_str(); }
int qmcb_host_vmc_create(qmcb_host_vmc** d, qmcb_crowd** crowds, const int* nw_per_crowd, int ncrowds, int n_electrons,
                         int precision, const uint32_t* seeds, double tau, int use_drift)
{
  try
  {
    if (!d || !crowds || !nw_per_crowd || !seeds || ncrowds <= 0)
      throw std::runtime_error("qmcb_host_vmc_create: bad arguments");
    auto* h       = new qmcb_host_vmc;
    h->N          = n_electrons;
    h->precision  = precision;
    h->tau        = tau;
    h->use_drift  = use_drift != 0;
    int off       = 0;
    for (int c = 0; c < ncrowds; ++c)
    {
      h->crowds.emplace_back(crowds[c], nw_per_crowd[c], off, seeds[c]);
      off += nw_per_crowd[c];
    }
    h->nw_total = off;
    *d          = h;
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
}
int qmcb_host_vmc_destroy(qmcb_host_vmc* d)
{
  delete d;
  return 0;
}
int qmcb_host_vmc_run(qmcb_host_vmc* d, int nsteps, uint8_t* accept_log_host)
{
  if (!d)
    return 1;
  std::atomic<int> failed{0};
  std::string first_error;
  for (int step = 0; step < nsteps && !failed; ++step)
  {
    uint8_t* log = accept_log_host ? accept_log_host + (size_t)step * d->N * d->nw_total : nullptr;
    auto work    = [&](int c) {
      try
      {
        if (d->precision == QMCB_MIXED)
          advance_crowd<float>(d->crowds[c], d->N, d->tau, d->use_drift, log, d->nw_total);
        else
          advance_crowd<double>(d->crowds[c], d->N, d->tau, d->use_drift, log, d->nw_total);
      }
      catch (const std::exception& e)
      {
        if (!failed.exchange(1))
          first_error = e.what();
      }
    };
    if (d->crowds.size() == 1)
      work(0);
    else
    {
      std::vector<std::thread> th;
      for (int c = 0; c < (int)d->crowds.size(); ++c)
        th.emplace_back(work, c);
      for (auto& t : th)
        t.join();
    }
  }
  if (failed)
  {
    g_err = first_error;
    return 1;
  }
  return 0;
}
int qmcb_host_vmc_counts(qmcb_host_vmc* d, long long* n_accept, long long* n_reject)
{
  long long a = 0, r = 0;
  for (auto& c : d->crowds)
  {
    a += c.n_accept;
    r += c.n_reject;
  }
  *n_accept = a;
  *n_reject = r;
  return 0;
}
int qmcb_host_vmc_bytes_per_sweep(qmcb_host_vmc* d, long long* h2d, long long* d2h)
{
  const long long T = d->precision == QMCB_MIXED ? 4 : 8;
  long long up = 0, down = 0;
  for (auto& c : d->crowds)
  {
    // per move: displacement [nw][3] T + accept flags [nw] up; grads_now [nw][3] T (drift only) + ratios [nw] f64 + grads_new [nw][3] T down
    const long long cs = qmcb_crowd_is_complex(c.crowd) ? 2 : 1; // complex gradients / ratios are (re, im) pairs
    up += (long long)d->N * ((long long)c.nw * 3 * T + c.nw);
    down += (long long)d->N * ((d->use_drift ? (long long)c.nw * 3 * T * cs : 0) + (long long)c.nw * 8 * cs + (long long)c.nw * 3 * T * cs);
  }
  *h2d = up;
  *d2h = down;
  return 0;
}
}
