// integration/smoke_main.cpp -- a C++ caller of libqmcb.so through include/qmcb.h only (no Python, no torch): what a
// QMCPACK driver thread does with one crowd, written out against the C ABI.
//
//   build:  integration/check.sh          (g++, links qmcpack_b200/libqmcb.so)
//   run:    integration/_build/qmcb_smoke  (needs a B200; without a CUDA device qmcb_init fails loudly and the
//                                           program exits with code 2 -- there is no CPU fallback)
//
// Sequence (reference call sites in parentheses):
//   1. table -> HBM (SPOSet::finalizeConstruction), crowd creation (createResource), loadWalker positions
//   2. mw_recompute, mw_evaluateGL                               (TrialWaveFunction.cpp:869, DiracDeterminantBatched.cpp:1122)
//   3. one host-driven particle-by-particle sweep: mw_evalGrad / mw_makeMove / mw_calcRatioGrad / mw_accept_rejectMove
//      with the Metropolis test on the host                       (VMCBatched.cpp:106-176)
//   4. mw_completeUpdates; the accumulated log psi must equal log psi0 + sum log|ratio| of the accepted moves and the
//      from-scratch value after mw_recompute (the reference's "delayed update == fresh inverse" check,
//      test_DiracDeterminantBatched.cpp:262-470)
//   5. two sweeps of the device-resident driver (qmcb_vmc_sweep), same consistency check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "qmcb.h"

#define CHECK(call)                                                              \
  do                                                                             \
  {                                                                              \
    if ((call) != 0)                                                             \
    {                                                                            \
      std::fprintf(stderr, "qmcb_smoke: %s failed: %s\n", #call, qmcb_last_error()); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

namespace
{
// minimal LCG + Box-Muller (the harness only needs reproducible numbers, not QMCPACK's stream)
struct Lcg
{
  unsigned long long s;
  double uniform()
  {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)((s >> 11) & ((1ull << 53) - 1)) / 9007199254740992.0;
  }
  double gauss() { return std::sqrt(-2.0 * std::log(1.0 - uniform())) * std::cos(6.283185307179586 * uniform()); }
};

// smooth orbitals: component m = cos or sin of one of the lowest reciprocal vectors, mixed by a fixed rotation so that
// the Slater matrices are well conditioned; sampled on the coefficient grid (periodic, 3 wrap-around planes)
std::vector<double> make_table(int M, int n, size_t npad, unsigned seed)
{
  const int P = M + 3;
  std::vector<double> c((size_t)P * P * P * npad, 0.0);
  std::vector<int> g;
  for (int s = 0; (int)g.size() < 3 * n; ++s)
    for (int i = -2; i <= 2 && (int)g.size() < 3 * n; ++i)
      for (int j = -2; j <= 2 && (int)g.size() < 3 * n; ++j)
        for (int k = -2; k <= 2 && (int)g.size() < 3 * n; ++k)
          if (i * i + j * j + k * k == s)
          {
            g.push_back(i), g.push_back(j), g.push_back(k);
          }
  Lcg rng{seed};
  std::vector<double> mix((size_t)n * n);
  for (auto& v : mix)
    v = rng.gauss();
  for (int d = 0; d < n; ++d)
    mix[(size_t)d * n + d] += 3.0; // diagonally dominant: the mixed set stays linearly independent
  for (int ix = 0; ix < P; ++ix)
    for (int iy = 0; iy < P; ++iy)
      for (int iz = 0; iz < P; ++iz)
      {
        double basis[64];
        for (int b = 0; b < n; ++b)
        {
          const double ph = 6.283185307179586 * (g[3 * b] * (double)ix + g[3 * b + 1] * (double)iy + g[3 * b + 2] * (double)iz) / M;
          basis[b]        = (b & 1) ? std::sin(ph) + 0.3 * std::cos(2 * ph) : std::cos(ph);
        }
        double* row = &c[(((size_t)ix * P + iy) * P + iz) * npad];
        for (int m = 0; m < n; ++m)
        {
          double v = 0;
          for (int b = 0; b < n; ++b)
            v += mix[(size_t)m * n + b] * basis[b];
          row[m] = v;
        }
      }
  return c;
}
} // namespace

int main()
{
  if (qmcb_init(0) != 0)
  {
    std::fprintf(stderr, "qmcb_smoke: no usable B200: %s\n", qmcb_last_error());
    return 2;
  }
  const int n = 8, N = 2 * n, nw = 6, M = 10, k = 4;
  const double L = 6.0;
  const size_t npad = qmcb_aligned_size(QMCB_FULL, n);
  const int grid[3] = {M, M, M};
  const double G[9] = {1 / L, 0, 0, 0, 1 / L, 0, 0, 0, 1 / L};
  qmcb_spline* spo[2];
  for (int s = 0; s < 2; ++s)
  {
    const std::vector<double> coefs = make_table(M, n, npad, 11u + s);
    CHECK(qmcb_spline_create(&spo[s], QMCB_FULL, QMCB_R2R, grid, n, n, npad, coefs.data(), G, nullptr, nullptr));
  }
  const double j2_uu[6] = {0.28, 0.19, 0.13, 0.08, 0.04, 0.01}, j2_ud[6] = {0.36, 0.22, 0.14, 0.09, 0.05, 0.02};
  const double ion_pos[6] = {1.5, 1.5, 1.5, 4.5, 4.5, 4.5};
  const int ion_grp[2]    = {0, 1};
  const double j1_params[12] = {-0.22, -0.18, -0.11, -0.04, 0.01, 0.0, -0.8, -0.6, -0.4, -0.2, -0.1, 0.0};
  const double j1_rcut[2]    = {2.9, 2.9};
  qmcb_system sys{};
  sys.precision = QMCB_FULL;
  sys.n_up = sys.n_dn = n;
  for (int i = 0; i < 3; ++i)
    sys.lattice[4 * i] = L;
  sys.spo[0] = spo[0], sys.spo[1] = spo[1];
  sys.delay_rank = k;
  sys.n_j2 = 6, sys.j2_uu = j2_uu, sys.j2_ud = j2_ud, sys.j2_rcut = 2.9;
  sys.nions = 2, sys.ion_pos = ion_pos, sys.ion_grp = ion_grp, sys.n_ion_groups = 2;
  sys.n_j1 = 6, sys.j1_params = j1_params, sys.j1_rcut = j1_rcut;
  qmcb_crowd* crowd;
  CHECK(qmcb_crowd_create(&crowd, &sys, nw));

  Lcg rng{2024};
  std::vector<double> R((size_t)nw * N * 3);
  for (auto& x : R)
    x = L * rng.uniform();
  CHECK(qmcb_crowd_set_positions(crowd, R.data()));
  CHECK(qmcb_twf_mw_recompute(crowd));
  std::vector<double> logpsi0(nw), logpsi1(nw), logpsi2(nw), ke(nw);
  CHECK(qmcb_twf_mw_evaluate_gl(crowd, nullptr, nullptr, logpsi0.data(), ke.data()));

  // ---- host-driven sweep
  const double tau = 0.05, sqrttau = std::sqrt(tau);
  std::vector<double> grad(3 * nw), gradnew(3 * nw), displ(3 * nw), ratio(nw), sumlog(nw, 0.0);
  std::vector<uint8_t> acc(nw);
  long long n_acc = 0;
  for (int iat = 0; iat < N; ++iat)
  {
    CHECK(qmcb_twf_mw_eval_grad(crowd, iat, grad.data()));
    for (int i = 0; i < 3 * nw; ++i)
      displ[i] = tau * grad[i] * 0.1 + sqrttau * rng.gauss(); // (a damped drift: enough to exercise the gradient path)
    CHECK(qmcb_ps_mw_make_move(crowd, iat, displ.data()));
    CHECK(qmcb_twf_mw_calc_ratio_grad(crowd, iat, ratio.data(), gradnew.data()));
    for (int iw = 0; iw < nw; ++iw)
    {
      if (!std::isfinite(ratio[iw]))
      {
        std::fprintf(stderr, "qmcb_smoke: non-finite ratio (walker %d, electron %d)\n", iw, iat);
        return 1;
      }
      acc[iw] = ratio[iw] * ratio[iw] > rng.uniform() ? 1 : 0;
      if (acc[iw])
        sumlog[iw] += std::log(std::fabs(ratio[iw])), ++n_acc;
    }
    CHECK(qmcb_twf_mw_accept_reject(crowd, iat, acc.data(), 1));
  }
  CHECK(qmcb_twf_mw_complete_updates(crowd));
  CHECK(qmcb_twf_mw_evaluate_gl(crowd, nullptr, nullptr, logpsi1.data(), ke.data()));
  CHECK(qmcb_twf_mw_recompute(crowd));
  CHECK(qmcb_twf_mw_evaluate_gl(crowd, nullptr, nullptr, logpsi2.data(), ke.data()));
  double worst = 0;
  for (int iw = 0; iw < nw; ++iw)
  {
    worst = std::fmax(worst, std::fabs(logpsi1[iw] - logpsi2[iw]));
    worst = std::fmax(worst, std::fabs(logpsi0[iw] + sumlog[iw] - logpsi2[iw]));
  }
  std::printf("host-driven sweep: %lld of %d moves accepted, |log psi (delayed) - log psi (recomputed)| <= %.2e, mode %d\n",
              n_acc, nw * N, worst, qmcb_crowd_host_kernel(crowd));
  if (!(worst < 1e-8) || n_acc == 0)
  {
    std::fprintf(stderr, "qmcb_smoke: delayed-update state disagrees with the from-scratch evaluation\n");
    return 1;
  }

  // ---- device-resident driver
  qmcb_vmc_params vp{};
  vp.tau = tau, vp.use_drift = 1, vp.seed = 7, vp.use_cuda_graph = 1;
  CHECK(qmcb_vmc_init(crowd, &vp));
  CHECK(qmcb_vmc_sweep(crowd, 2, nullptr));
  std::vector<long long> nacc(nw), nrej(nw); // per walker
  CHECK(qmcb_vmc_counts(crowd, nacc.data(), nrej.data()));
  long long na = 0, nr = 0;
  for (int iw = 0; iw < nw; ++iw)
    na += nacc[iw], nr += nrej[iw];
  CHECK(qmcb_twf_mw_evaluate_gl(crowd, nullptr, nullptr, logpsi1.data(), ke.data()));
  CHECK(qmcb_twf_mw_recompute(crowd));
  CHECK(qmcb_twf_mw_evaluate_gl(crowd, nullptr, nullptr, logpsi2.data(), ke.data()));
  worst = 0;
  for (int iw = 0; iw < nw; ++iw)
    worst = std::fmax(worst, std::fabs(logpsi1[iw] - logpsi2[iw]));
  std::printf("device sweeps: %lld accepted, %lld rejected (sweep kernel %d), log psi drift vs recompute %.2e, %llu launches\n",
              na, nr, qmcb_vmc_sweep_kernel(crowd), worst, qmcb_kernel_launch_count());
  if (!(worst < 1e-8) || na + nr != 2ll * nw * N || na == 0)
  {
    std::fprintf(stderr, "qmcb_smoke: device sweep inconsistent\n");
    return 1;
  }
  CHECK(qmcb_crowd_destroy(crowd));
  CHECK(qmcb_spline_destroy(spo[0]));
  CHECK(qmcb_spline_destroy(spo[1]));
  std::printf("qmcb_smoke ok\n");
  return 0;
}
