// qmcpack_b200/csrc/driver.cuh -- device-resident pieces of the particle-by-particle VMC sweep (sm_100a).
//
// Restates on the device what VMCBatched::advanceWalkers does on the host between the mw_* calls
// (QMCDrivers/VMC/VMCBatched.cpp:106-176) so that a whole sweep needs no host round trip:
//   * std::mt19937 raw stream (Utilities/StdRandom.h:59) produced by one CTA into a ring buffer
//   * uniform = raw / 2^32 in double (StdRandom.h:43-47); Box-Muller pairs (Particle/ParticleBase/RandomSeqGenerator.h:33-52)
//   * drift = UNR-scaled gradient (GreenFunctionModifiers/DriftModifierUNR.cpp:20-31), log Green's functions
//     (QMCDriverNew.h:282-290), acceptance test with the uniform drawn ONLY when the move is valid and
//     prob >= eps, walkers in index order (VMCBatched.cpp:152-167) -> exclusive scan of the "needs a draw" flags
#pragma once
#include "common.cuh"

namespace qmcb
{
struct RngDev
{
  uint32_t* state;          // [624] mt19937 state vector
  uint32_t* ring;           // [ring_size] raw outputs
  unsigned long long* gen;  // number of raw outputs generated so far
  unsigned long long* pos;  // pos[2]: number consumed so far, double-buffered by move parity: the Metropolis test of
                            // electron iat reads pos[iat & 1] and publishes pos[(iat + 1) & 1], so CTAs that start late
                            // never see the value their own kernel wrote; the sweep prologue leaves its result in pos[0]
  unsigned* sweep;          // sweeps started so far (epoch of the cross-CTA flags)
  unsigned* flags;          // [nw] (epoch << 1 | needs_draw) published by each walker's CTA
  unsigned ring_mask;       // ring_size - 1 (power of two)
};

template<typename RT>
struct DriverDev
{
  int nw, N;
  RT tauovermass, oneover2tau, sqrttau;
  int use_drift;
  int pdl_early;   // signal programmatic dependents at kernel start
  int l1_prefetch; // boundary kernel pulls the Jastrow inputs into L1 at its start (env QMCB_L1PF, default on)
  int fuse_jastrow; // device sweep: the proposed move's Jastrow sums run in the boundary kernel's tail
  int dmc;         // DMCBatched acceptance rule + rr accumulators (DMCBatched.cpp:188-250)
  RT* rr_accepted; // [nw] sum over accepted moves of tau |delta|^2 (this sweep)
  RT* rr_proposed; // [nw] the same over all proposed moves
  RT* deltas;      // [N][nw][3] Gaussians of the current sweep (walker_deltas, VMCBatched.cpp:109,122)
  RT* drifts;      // [nw][3]   displacement actually proposed (drift + delta)
  RT* delta_cur;   // [nw][3]
  RT* grads_now;   // [nw][3] component sum
  unsigned char* accepted; // [nw]
  unsigned long long* n_accept; // [nw]
  unsigned long long* n_reject; // [nw]
  unsigned char* accept_log;    // optional [N][nw] of the current sweep
  unsigned* err;                // crowd-wide error bits (det.cuh: QMCB_ERR_*)
  int value_only;               // the proposed move's orbital rows hold VALUES only (TrialWaveFunction::mw_calcRatio):
                                // an accept must not touch the gradient / Laplacian rows
  unsigned* ticket;             // not null: the boundary kernel's CTAs take their walker index from this counter in the
                                // order they START instead of from blockIdx.x.  The Metropolis warp of walker iw waits for
                                // the flags of the walkers below it (RNG order); when a launch has more CTAs than the
                                // device can hold at once that is only deadlock-free if every lower-index walker is owned
                                // by a CTA that is already running -- which the dispatch-order ticket guarantees and the
                                // hardware's block scheduling order does not promise.  Set by the host for oversubscribed
                                // launches only (and by QMCB_TICKET=1).
};

#ifdef __CUDACC__
// ---- std::mt19937 -------------------------------------------------------------------------------
__global__ void mt19937_seed_kernel(RngDev R, uint32_t seed)
{
  if (threadIdx.x == 0 && blockIdx.x == 0)
  {
    // [rand.eng.mers] seeding: x_i = 1812433253 * (x_{i-1} ^ (x_{i-1} >> 30)) + i
    uint32_t x = seed;
    R.state[0] = x;
    for (int i = 1; i < 624; ++i)
    {
      x          = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
      R.state[i] = x;
    }
    *R.gen   = 0;
    R.pos[0] = R.pos[1] = 0;
    *R.sweep = 0;
  }
}

// tops the ring up until (gen - pos) >= backlog.  One CTA of 256 threads; the 624-word twist runs in 4 dependent
// phases ([0,227) uses old words only, [227,454) and [454,623) use words produced one phase earlier, word 623 last).
__global__ void __launch_bounds__(256) mt19937_fill_kernel(RngDev R, unsigned long long backlog)
{
  __shared__ uint32_t x[624];
  __shared__ unsigned long long s_gen;
  const int tid = threadIdx.x;
  for (int i = tid; i < 624; i += 256)
    x[i] = R.state[i];
  if (tid == 0)
    s_gen = *R.gen;
  __syncthreads();
  // consumption counter: the two parity slots leapfrog (driver.cuh RngDev::pos), the larger one is the more recent;
  // possibly stale (smaller): see the fork comment in enqueue_sweep
  const unsigned long long pos = R.pos[0] > R.pos[1] ? R.pos[0] : R.pos[1];
  unsigned long long gen       = s_gen;
  auto twist = [](uint32_t xi, uint32_t xi1, uint32_t xm) {
    const uint32_t y = (xi & 0x80000000u) | (xi1 & 0x7fffffffu);
    return xm ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  };
  while (gen - pos < backlog)
  {
    uint32_t v;
    // phase A: i in [0,227)
    if (tid < 227)
      v = twist(x[tid], x[tid + 1], x[tid + 397]);
    __syncthreads();
    if (tid < 227)
      x[tid] = v;
    __syncthreads();
    // phase B: i in [227,454)
    if (tid < 227)
      v = twist(x[tid + 227], x[tid + 228], x[tid]);
    __syncthreads();
    if (tid < 227)
      x[tid + 227] = v;
    __syncthreads();
    // phase C: i in [454,623)
    if (tid < 169)
      v = twist(x[tid + 454], x[tid + 455], x[tid + 227]);
    __syncthreads();
    if (tid < 169)
      x[tid + 454] = v;
    __syncthreads();
    if (tid == 0)
      x[623] = twist(x[623], x[0], x[396]);
    __syncthreads();
    // tempering
    for (int i = tid; i < 624; i += 256)
    {
      uint32_t y = x[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      R.ring[(unsigned)((gen + i) & R.ring_mask)] = y;
    }
    gen += 624;
    __syncthreads();
  }
  for (int i = tid; i < 624; i += 256)
    R.state[i] = x[i];
  if (tid == 0)
    *R.gen = gen;
}

__device__ __forceinline__ double rng_uniform(const RngDev& R, unsigned long long idx)
{
  return (double)R.ring[(unsigned)(idx & R.ring_mask)] / 4294967296.0;
}

// Gaussians of one sub-step: count = 3*nw*N values from pairs of uniforms starting at *pos; the kernel that runs last
// in stream order (gauss_advance_kernel) moves *pos.
template<typename RT>
__global__ void gauss_kernel(RngDev R, RT* out, unsigned long long count, int last_parity)
{
  const unsigned long long pair = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long i    = 2 * pair;
  if (i >= count)
    return;
  const unsigned long long base   = R.pos[last_parity]; // left by the last Metropolis test of the previous sweep
  const double slightly_less_one  = 1.0 - 2.220446049250313e-16;
  const double u1 = rng_uniform(R, base + i), u2 = rng_uniform(R, base + i + 1);
  const double t1 = sqrt(-2.0 * log(1.0 - slightly_less_one * u1));
  const double t2 = 2.0 * 3.14159265358979323846 * u2;
  double s, c;
  sincos(t2, &s, &c);
  out[i] = (RT)(t1 * c);
  if (i + 1 < count)
    out[i + 1] = (RT)(t1 * s);
}
__global__ void rng_advance_kernel(RngDev R, unsigned long long n, int last_parity)
{
  if (threadIdx.x == 0 && blockIdx.x == 0)
  {
    R.pos[0] = R.pos[last_parity] + n;
    *R.sweep += 1;
  }
}

// ref: DriftModifierUNR.cpp:20-31 (a = 1)
// (QMCB_SEG_SMALL: one out-of-line copy instead of one inlined copy per caller -- a double-precision square root and
// division each; the walker-segment kernel calls it from four places and is instruction-cache bound, segment.cuh)
#ifndef QMCB_SEG_SMALL
#define QMCB_SEG_SMALL 1
#endif
#if QMCB_SEG_SMALL
#define QMCB_DRIFT_INLINE __noinline__
#else
#define QMCB_DRIFT_INLINE __forceinline__
#endif
template<typename RT>
__device__ QMCB_DRIFT_INLINE RT drift_scale(const RT tau, const RT vsq)
{
  const RT eps = sizeof(RT) == 4 ? RT(1.1920929e-07f) : RT(2.220446049250313e-16);
  return vsq < eps ? tau : (RT)((-1.0 + sqrt(1.0 + 2.0 * (double)RT(1) * (double)tau * (double)vsq)) / (double)(RT(1) * vsq));
}
template<typename RT>
__device__ __forceinline__ void get_drift(RT tau, const RT qf[3], RT drift[3])
{
  const RT vsq = qf[0] * qf[0] + qf[1] * qf[1] + qf[2] * qf[2];
  const RT sc  = drift_scale<RT>(tau, vsq);
  drift[0]     = qf[0] * sc;
  drift[1]     = qf[1] * sc;
  drift[2]     = qf[2] * sc;
}
#endif

} // namespace qmcb
