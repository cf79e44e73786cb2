// qmcpack_b200/csrc/segment.cuh -- the persistent walker-segment kernel of the device-resident sweep (sm_100a).
//
// One CTA owns ONE walker for a whole segment of the particle-by-particle loop (the <= delay_rank electron moves between
// two Woodbury flushes) and runs, per move, everything the reference spreads over TrialWaveFunction::mw_evalGrad /
// mw_calcRatioGrad / mw_accept_rejectMove and VMCBatched::advanceWalkers (QMCDrivers/VMC/VMCBatched.cpp:106-176):
//
//   propose (UNR drift + Gaussian)                                          DriftModifierUNR.cpp:20-31, ParticleSet::mw_makeMove
//   spline gather of the 4x4x4 stencil + VGL + det ratio / gradient dots     SplineR2R.cpp:414-582 (mw_evaluateVGLandDetRatioGrads)
//   J1 / J2 sums at the proposed position                                    TwoBodyJastrow.cpp:542-578, J1OrbitalSoA ratioGrad
//   Metropolis test in the crowd's RNG order                                 VMCBatched.cpp:139-167 (DMC rule DMCBatched.cpp:188-250)
//   determinant accept / pseudo-accept, Jastrow accept, position commit      DelayedUpdateBatched.h:542-670, TwoBodyJastrow.cpp:631-664
//   inverse row of the next electron incl. pending delays + old gradient     DelayedUpdateBatched.h:174-235, :354-400
//
// Why one kernel: with one launch per step (move_boundary_kernel + spline_gather_kernel, crowd.cu / spline.cuh) every
// move pays two launch ramps and all walkers march in lock step -- the HBM stream of the gather idles while the
// boundary's chain of dependent L2 round trips runs and vice versa.  Here every walker's CTA is resident for the whole
// segment (4 CTAs per SM at NiO-a64 in mixed precision = 592 slots for 512 walkers), walkers drift apart by up to one
// move, and one walker's stencil streams from HBM while its SM neighbours sit in their boundary chains.  Binv, w, the
// proposed orbital rows and the partial dots never leave shared memory.
//
// Stencil staging: the table is the same 4-D TMA tensor (component, z, y, x) as in spline.cuh, but a stage is a SLAB of
// the stencil -- all components x 4 z x 2 y x 1 x = 8 rows -- so that the 64 rows of an evaluation stream through a
// 3-stage ring of 8-row slabs (36 KB at 384 float components) in the reference's accumulation order (x index outer, y
// index inner, MultiBsplineVGLH.hpp:164-202); the accumulators stay in registers across the slabs.
//
// Cross-walker order: the reference draws the uniform of a move only when prob >= eps, walkers in index order, moves in
// electron order (VMCBatched.cpp:156-158).  Walker iw's stream position at move m is
//     tot[m] + #{ j < iw : need[m][j] },      tot[m + 1] = tot[m] + #{ all j : need[m][j] }.
// Every CTA publishes need[m][iw] (tagged with the sweep number) before it looks back, the CTA of the last walker
// publishes tot[m + 1]; nobody waits for a Metropolis DECISION of another walker.  All CTAs of the launch must be
// co-resident (the host checks the occupancy and reserves the slots per device, crowd.cu).
#pragma once
#include "spline.cuh"
#include "det.cuh"
#include "jastrow.cuh"
#include "driver.cuh"

namespace qmcb
{
constexpr int SEG_TPB    = 256; // warps 0-5: spline consumers; warp 6: TMA producer; warp 7: Jastrow sums + Metropolis
constexpr int SEG_NCONS  = 192;
constexpr int SEG_NSTAGE = 3;
constexpr int SEG_ROWS   = 8;   // stencil rows per stage
constexpr int SEG_BOXW   = 192; // components per TMA box (two boxes per stage at CPT = 2)
constexpr int SEG_NQ     = 8;   // stages per evaluation
// warps that share the Jastrow sums of the proposed position: the six consumers (6), or the consumers plus warp 7 (7), which
// is idle under the gather once its Metropolis inputs are fetched (832 candidates at NiO-a64: 4 blocks per lane instead of
// 5 on the critical warps).  Measured: 31.3 ms of segment kernels per sweep with 7 against 28.5 ms with 6.
#ifndef QMCB_SEG_JW
#define QMCB_SEG_JW 6
#endif
constexpr int SEG_JW = QMCB_SEG_JW;
// Threads of the determinant accept + next-row group in the phase after the Metropolis decision; the other SEG_TPB - SEG_DET
// run the Jastrow accept.  An absolute-clock trace of the phase (QMCB_SEG_TRACE) shows the determinant side done in ~4 us
// on 128 threads while the Jastrow accept of an accepted move takes ~12 us on the other 128 and holds the CTA barrier;
// giving the Jastrow side 160 or 192 threads changed nothing (30.8 / 31.5 / 30.7 ms of segment kernels per sweep at
// 64 / 96 / 128 determinant threads): its time is a chain of dependent accesses, not a lack of threads.  A split that
// follows the decision (64 determinant threads on an accepted move, 128 on a rejected one) was slower too: 29.3 vs 28.5 ms.
#ifndef QMCB_SEG_DET
#define QMCB_SEG_DET 128
#endif
constexpr int SEG_DET = QMCB_SEG_DET;

// per-sweep arrays of the cross-walker RNG order (zero-initialised once; entries are tagged with the sweep number)
struct SegRng
{
  unsigned* flags;             // [N][stride]  (tag << 1) | needs_draw
  unsigned long long* tot_val; // [N + 1] raw outputs consumed before move m
  unsigned* tot_tag;           // [N + 1]
  unsigned* btot;              // [N][32]  (tag << 8) | draws needed by the walkers 32 b .. 32 b + 31 (published by the block's last walker)
  int stride;                  // walker capacity of the crowd
};

struct SegLayout
{
  unsigned ring, phi, x, Bs, vec, red, rgp, sg, hdr, bars, jred, jsum, tim, jl, total;
  int jl_entries;
};

// shared-memory carve-up (bytes).  CPT = components per consumer thread (1: n <= 192, 2: n <= 384)
template<typename T, int CPT>
__host__ __device__ inline SegLayout seg_layout(int n, int k, int N, int nions)
{
  auto up = [](unsigned v, unsigned a) { return (v + a - 1) / a * a; };
  SegLayout L;
  unsigned o = 0;
  L.ring = o, o += SEG_NSTAGE * SEG_ROWS * CPT * SEG_BOXW * (unsigned)sizeof(T);
  L.phi = o, o += up(5 * n * (unsigned)sizeof(T), 16);
  L.x = o, o += up(n * (unsigned)sizeof(T), 16);
  L.Bs = o, o += up(k * (k + 1) * (unsigned)sizeof(T), 16);
  L.vec = o, o += up(4 * k * (unsigned)sizeof(T), 16); // pA, pB, y, w
  L.red = o, o += 96 * (unsigned)sizeof(T);
  L.rgp = o, o += 32 * (unsigned)sizeof(T);
  L.sg = o, o += 16 * (unsigned)sizeof(T);
  L.hdr = o, o += up((SPL_HDR + SPL_SCRATCH + 4) * (unsigned)sizeof(T), 16);
  L.bars = o, o += 64;
  L.jred = o, o += 128 * (unsigned)sizeof(T); // per-warp partials of the ten Jastrow sums [SEG_JW][16]
  L.jsum = o, o += 16 * (unsigned)sizeof(T); // the sums of the move being decided
  L.tim = o = up(o, 16), o += 192 + 16;      // cycle stamps of the timing build; the last 16 bytes: log|det|, phase
  L.jl_entries = (int)up((unsigned)(N + nions + 256), 32);
  L.jl = o, o += up(2u * (unsigned)L.jl_entries, 16);
  L.total = o;
  return L;
}

#ifdef __CUDACC__
// per-phase cycle stamps of the walker-segment kernel (debug builds only: scripts/build_variant.sh -DQMCB_SEG_TIMING)
#ifdef QMCB_SEG_TIMING
#define SEG_STAMP(slot, who)                                             \
  do                                                                     \
  {                                                                      \
    if (threadIdx.x == (who))                                            \
    {                                                                    \
      const long long now__ = clock64();                                 \
      seg_tim[slot] += now__ - seg_tim[16 + ((who) >> 5)];               \
      seg_tim[16 + ((who) >> 5)] = now__;                                \
    }                                                                    \
  } while (0)
#else
#define SEG_STAMP(slot, who) \
  do                         \
  {                          \
  } while (0)
#endif
// absolute-clock trace of ONE walker's CTA over a few moves (debug builds: -DQMCB_SEG_TRACE=<walker index>): lane 0 of every
// warp records clock64() at labelled points into a global array that the CTA prints at the end of the launch -- the only way
// to see the critical path across the warp roles and barriers of the kernel
#ifdef QMCB_SEG_TRACE
__device__ long long g_seg_trace[4][8][40];
#define SEG_TR(id)                                                                                   \
  do                                                                                                 \
  {                                                                                                  \
    if (trace_on && (threadIdx.x & 31) == 0)                                                         \
      g_seg_trace[m - 8][threadIdx.x >> 5][id] = clock64();                                          \
  } while (0)
#else
#define SEG_TR(id) \
  do               \
  {                \
  } while (0)
#endif
namespace ptx
{
__device__ __forceinline__ void fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
} // namespace ptx

template<typename T, int CPT>
struct SegVec;
template<typename T>
struct SegVec<T, 1>
{
  static __device__ __forceinline__ void load(const T* p, T (&o)[1]) { o[0] = p[0]; }
};
template<>
struct SegVec<float, 2>
{
  static __device__ __forceinline__ void load(const float* p, float (&o)[2])
  {
    const float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = v.x, o[1] = v.y;
  }
};
template<>
struct SegVec<double, 2>
{
  static __device__ __forceinline__ void load(const double* p, double (&o)[2])
  {
    const double2 v = *reinterpret_cast<const double2*>(p);
    o[0] = v.x, o[1] = v.y;
  }
};

// four consecutive values with 16-byte loads (rows whose byte length is a multiple of 16)
template<typename T>
struct V4
{
  T x, y, z, w;
};
__device__ __forceinline__ V4<float> ld4(const float* p)
{
  const float4 v = *reinterpret_cast<const float4*>(p);
  return V4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ V4<double> ld4(const double* p)
{
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  return V4<double>{a.x, a.y, b.x, b.y};
}
template<typename T>
__device__ __forceinline__ T dot4(const V4<T>& a, const V4<T>& b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// proposal of electron iat for walker iw by one warp: lanes 0-2 own one Cartesian component each; every lane returns
// the proposed position.  g_det: determinant gradient of the prepared row (shared memory).  STORE: publish the
// displacement, the Gaussian part and the position (Metropolis test, Jastrow accept and API readers take them from memory)
template<typename T, bool STORE>
__device__ __forceinline__ void seg_propose(const DriverDev<T>& Dr, const JastrowDev<T>& J, const int iw, const int iat,
                                            const T* g_det, T newpos[3], T* disp_out = nullptr, T* delta_out = nullptr)
{
  const int lane = threadIdx.x & 31;
  const int d    = lane < 3 ? lane : 0;
  const T delta  = Dr.deltas[((size_t)iat * Dr.nw + iw) * 3 + d] * Dr.sqrttau;
  const T rold   = J.rsoa[((size_t)iw * 3 + d) * J.npad + iat];
  T disp         = delta;
  if (Dr.use_drift)
  {
    T gd = g_det[d];
    if (J.has_j2)
      gd += J.dUat[((size_t)iw * 3 + d) * J.npad + iat];
    if (J.has_j1)
      gd += J.Grad1[((size_t)iw * 3 + d) * J.N + iat];
    const T gv[3] = {__shfl_sync(0xffffffffu, gd, 0), __shfl_sync(0xffffffffu, gd, 1), __shfl_sync(0xffffffffu, gd, 2)};
    T dr[3];
    get_drift<T>(Dr.tauovermass, gv, dr);
    disp = (d == 0 ? dr[0] : (d == 1 ? dr[1] : dr[2])) + delta;
  }
  const T p = rold + disp;
  if (STORE && lane < 3)
  {
    Dr.drifts[3 * iw + d]    = disp;
    Dr.delta_cur[3 * iw + d] = delta;
    J.newpos[3 * iw + d]     = p;
  }
  newpos[0] = __shfl_sync(0xffffffffu, p, 0);
  newpos[1] = __shfl_sync(0xffffffffu, p, 1);
  newpos[2] = __shfl_sync(0xffffffffu, p, 2);
  if (disp_out)
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
      disp_out[e]  = __shfl_sync(0xffffffffu, disp, e);
      delta_out[e] = __shfl_sync(0xffffffffu, delta, e);
    }
}

// what the Metropolis test of (iw, iat) needs from memory besides the orbital dots: fetched by the Metropolis warp WHILE
// the spline gather of the same move runs, so that the test itself is arithmetic plus the cross-walker look-back.
// The crowd's stream position (tot[m], published by the LAST walker when its test of the previous move is done) is NOT
// waited for here: this warp would hold the CTA barrier behind the gather, and with it the staging and the dot sweeps
// of the other warps, which do not depend on it (an absolute-clock trace showed the consumers idle at that barrier for
// 4 - 17 us per move); seg_metropolis picks it up after publishing this walker's flag.
template<typename T>
struct SegMetroPre
{
  T uat_old, vat_old;          // Uat[iat], Vat[iat] of the committed configuration
  T disp[3], delta[3];         // proposed displacement and its Gaussian part
  T np[3];                     // the proposed position
  T rr;                        // DMC: tau |Gaussian|^2
  unsigned tag;                // sweep tag of the look-back arrays
  unsigned long long base;     // raw outputs consumed before this move
  uint32_t raw_spec;           // ring word at base + iw (the draw of this walker when every lower walker draws)
};

template<typename T>
__device__ __forceinline__ SegMetroPre<T> seg_metro_prefetch(const DriverDev<T>& Dr, const JastrowDev<T>& J, const RngDev& R,
                                                             const SegRng& SR, const int iw, const int iat, const T* g_det)
{
  SegMetroPre<T> P;
  T np3[3];
  seg_propose<T, false>(Dr, J, iw, iat, g_det, np3, P.disp, P.delta);
  P.np[0] = np3[0], P.np[1] = np3[1], P.np[2] = np3[2];
  P.uat_old = J.has_j2 ? J.Uat[(size_t)iw * J.npad + iat] : T(0);
  P.vat_old = J.has_j1 ? J.Vat[(size_t)iw * J.N + iat] : T(0);
  P.rr      = T(0);
  if (Dr.dmc)
  {
    const T* dr = Dr.deltas + ((size_t)iat * Dr.nw + iw) * 3;
    P.rr        = Dr.tauovermass * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
  }
  P.tag      = ((*R.sweep) & 0x3fffffffu) + 1u;
  P.base     = 0;
  P.raw_spec = 0;
  return P;
}

// Metropolis test of (walker iw, electron iat) by one warp.  q[4]: determinant ratio and gradient dots (undivided);
// js[10]: Jastrow sums at the proposed position (shared memory); P: seg_metro_prefetch.  Returns the decision in every
// lane; rdet_out = determinant ratio.
template<typename T>
__device__ __forceinline__ bool seg_metropolis(const DriverDev<T>& Dr, const JastrowDev<T>& J, const RngDev& R, const SegRng& SR,
                                               const int iw, const int iat, const T q[4], const T* js, const SegMetroPre<T>& P,
                                               T& rdet_out, long long* tr = nullptr)
{
#define SEG_MTR(id)                      \
  do                                     \
  {                                      \
    if (tr && (threadIdx.x & 31) == 0)   \
      tr[id] = clock64();                \
  } while (0)
  const int lane = threadIdx.x & 31;
  const T rdet   = q[0];
  double ratio   = (double)rdet;
  T gn[3]        = {q[1] / rdet, q[2] / rdet, q[3] / rdet};
  // (one copy of the double-precision exponential for both Jastrow ratios, J2 first: the kernel is instruction-cache bound)
#pragma unroll 1
  for (int t = 0; t < 2; ++t)
    if (t == 0 ? J.has_j2 : J.has_j1)
    {
      ratio = ratio * exp((double)((t == 0 ? P.uat_old : P.vat_old) - js[5 * t]));
      gn[0] += js[5 * t + 1];
      gn[1] += js[5 * t + 2];
      gn[2] += js[5 * t + 3];
    }
  const T eps = sizeof(T) == 4 ? T(1.1920929e-07f) : T(2.220446049250313e-16);
  T prob      = (T)(ratio * ratio); // std::norm(ratio), VMCBatched.cpp:152
  if (isnan(ratio) && lane == 0 && Dr.err)
    atomicOr(Dr.err, QMCB_ERR_NAN_RATIO); // NaNguard::checkOneParticleRatio (TrialWaveFunction.cpp:549): the host throws
  bool need             = prob >= eps;
  const unsigned tag    = P.tag;
  volatile unsigned* fl = SR.flags + (size_t)iat * SR.stride;
  // VMC: whether this walker draws depends on the ratio alone -- the flag goes out BEFORE the Green's functions are
  // evaluated (every walker above this one waits for it)
  if (!Dr.dmc && lane == 0)
    fl[iw] = (tag << 1) | (need ? 1u : 0u);
  T log_gf = T(0), log_gb = T(0);
  if (Dr.use_drift)
  {
    log_gf = -Dr.oneover2tau * (P.delta[0] * P.delta[0] + P.delta[1] * P.delta[1] + P.delta[2] * P.delta[2]);
    T dr[3];
    get_drift<T>(Dr.tauovermass, gn, dr);
    dr[0] += P.disp[0];
    dr[1] += P.disp[1];
    dr[2] += P.disp[2];
    log_gb = -Dr.oneover2tau * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
  }
  if (Dr.dmc)
  {
    // DMCBatched.cpp:188-250 (see metropolis_warp in crowd.cu)
    const bool reject = !(ratio > 0.0);
    prob              = (T)(ratio * ratio * (double)exp(log_gb - log_gf));
    need              = !reject && prob >= eps;
    if (lane == 0)
      fl[iw] = (tag << 1) | (need ? 1u : 0u);
  }
  SEG_MTR(30);
  // ---- #{ j < iw : need[j] }.  Two levels: the last walker of every block of 32 publishes the block's count as soon as
  // the block's flags are in; a walker adds the counts of the blocks below its own (one load per lane) and the flags of
  // its own block below itself (one load per lane) -- two short dependent round trips to L2.  The flat loop over all lower
  // flags (iw / 32 dependent polls per lane) cost the last walkers of a 512-walker crowd 5 - 9 us per move, and the last
  // walker's test is what every other walker waits for (`base` below).
  unsigned cnt   = 0;
  const int nblk = (Dr.nw + 31) >> 5;
  const bool two_level = nblk <= 32;
  const int lb = iw >> 5, li = iw & 31;
  volatile unsigned* bt = SR.btot + (size_t)iat * 32;
  const unsigned tag24  = (tag & 0xffffffu) << 8;
  unsigned below        = 0;
  if (two_level)
  {
    unsigned f = 0;
    if (lane < li)
      while (((f = fl[32 * lb + lane]) >> 1) != tag)
        __nanosleep(60);
    below = __popc(__ballot_sync(0xffffffffu, lane < li && (f & 1u)));
    if ((li == 31 || iw == Dr.nw - 1) && lane == 0)
      bt[lb] = tag24 | (below + (need ? 1u : 0u));
  }
  // raw outputs consumed before this move: left by the sweep prologue (rng_advance_kernel) for the first electron,
  // otherwise published by the last walker after its test of the previous move
  unsigned long long base;
  if (iat == 0)
    base = R.pos[0];
  else
  {
    volatile unsigned* tt = SR.tot_tag + iat;
    while (*tt != tag)
      __nanosleep(100);
    __threadfence();
    base = *((volatile unsigned long long*)(SR.tot_val + iat));
  }
  if (two_level)
  {
    unsigned v = tag24;
    if (lane < lb)
      while (((v = bt[lane]) & 0xffffff00u) != tag24)
        __nanosleep(60);
    cnt = (lane < lb ? (v & 0xffu) : 0u) + (lane == 0 ? below : 0u);
  }
  else
    for (int j = lane; j < iw; j += 32)
    {
      unsigned f;
      while (((f = fl[j]) >> 1) != tag)
        __nanosleep(100);
      cnt += f & 1u;
    }
  SEG_MTR(31);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  bool acc = false;
  if (lane == 0)
  {
    if (iw == Dr.nw - 1)
    {
      const unsigned long long next = base + cnt + (need ? 1u : 0u);
      *((volatile unsigned long long*)(SR.tot_val + iat + 1)) = next;
      __threadfence();
      *((volatile unsigned*)(SR.tot_tag + iat + 1)) = tag;
      if (iat == Dr.N - 1)
        R.pos[(iat + 1) & 1] = next; // where the next sweep's Gaussians start (gauss_kernel reads pos[N & 1])
    }
    if (need)
    {
      const uint32_t raw = R.ring[(unsigned)((base + cnt) & R.ring_mask)];
      const double u     = (double)raw / 4294967296.0;
      acc                = Dr.dmc ? (u < (double)prob) : (u < (double)(prob * exp(log_gb - log_gf)));
    }
    SEG_MTR(32);
    Dr.accepted[iw] = acc ? 1 : 0;
    if (Dr.dmc)
    {
      Dr.rr_proposed[iw] += P.rr;
      if (acc)
        Dr.rr_accepted[iw] += P.rr;
    }
    if (acc)
      Dr.n_accept[iw] += 1;
    else
      Dr.n_reject[iw] += 1;
    if (Dr.accept_log)
      Dr.accept_log[(size_t)iat * Dr.nw + iw] = acc ? 1 : 0;
  }
  rdet_out = rdet;
  SEG_MTR(33);
  return __shfl_sync(0xffffffffu, acc ? 1 : 0, 0) != 0;
#undef SEG_MTR
}

// ---- host-driven mode (template parameter HD): the Metropolis test and the drift stay with the CALLER (QMCPACK's batched
// driver calling TrialWaveFunction::mw_evalGrad / mw_calcRatioGrad / mw_accept_rejectMove per electron), but the kernel
// stays resident for the whole segment and talks to the host thread through mailboxes in pinned host memory instead of
// being launched four times per move:
//     kernel: gradient of the prepared electron -> h_out1                           (qmcb_twf_mw_eval_grad reads it)
//     host:   displacements -> h_in1                                                 (qmcb_ps_mw_make_move)
//     kernel: ratio, new gradient -> h_out2                                          (qmcb_twf_mw_calc_ratio_grad)
//     host:   accept flags -> h_in2                                                  (qmcb_twf_mw_accept_reject)
// Every message is made of 16-byte chunks that carry the sequence number of the exchange NEXT TO the payload and are
// written with one 16-byte store, so a chunk is either old or complete: no flag word behind the data, no system-scope
// fence between them, and the reader's polling load IS the data load (a PCIe read is a ~2 us round trip; "flag, then
// data" would pay it twice).  Only the CTA of walker 0 polls host memory; it copies the host's reply into device memory
// and republishes the sequence number there for the other CTAs.  h_cmd[0] = SEG_ABORT makes every CTA leave at its
// next wait with the state of the completed moves written back (the host falls back to the launch-per-call path, e.g.
// for an API call that is not part of the move loop).
constexpr unsigned SEG_ABORT = 0xffffffffu;
template<typename T>
struct SegHost
{
  uint4* h_out1;                // [nw][3]  {value (T, low bytes of 8), -, tag, -}: gradient components           pinned host
  uint4* h_out2;                // [nw][4]  chunk 0: ratio (double); chunks 1-3: new gradient components (T); tag in .z
  const uint4* h_in1;           // float: [nw] {dx, dy, dz, tag}; double: [nw][3] {value (8 bytes), tag, -}
  const uint4* h_in2;           // [ceil(nw / 4)] words (seq << 1) | accepted, one per walker
  volatile unsigned* h_cmd;     // [0] SEG_ABORT request (host -> kernel), [1] error word written by the kernel
  T* d_displ;                   // [nw][3] device copies published by walker 0's CTA
  unsigned char* d_acc;         // [nw]
  volatile unsigned* d_cmd;     // [1] last sequence number republished on the device, [1] moves completed by this launch
  unsigned seq0;                // sequence number of the first exchange of this launch
};

__device__ __forceinline__ uint4 ld_sys16(const uint4* p)
{
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_sys16(uint4* p, const uint4 v)
{
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// one outgoing chunk: the value's bits in .x/.y, the tag in .z
__device__ __forceinline__ void seg_post(uint4* slot, const float v, const unsigned tag)
{
  st_sys16(slot, make_uint4(__float_as_uint(v), 0u, tag, 0u));
}
__device__ __forceinline__ void seg_post(uint4* slot, const double v, const unsigned tag)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  st_sys16(slot, make_uint4((unsigned)b, (unsigned)(b >> 32), tag, 0u));
}

// One warp waits until the host has answered exchange `seq` (KIND 1: displacements, KIND 2: accept flags); the warp of
// walker 0 polls the tagged chunks in host memory -- all loads of a pass in flight at once, up to 8 per lane -- unpacks
// them into device memory and republishes.  Returns false on abort / time-out.
template<typename T, int KIND>
__device__ __forceinline__ bool seg_host_wait(const SegHost<T>& H, const int iw, const int nw, const unsigned seq)
{
  const int lane = threadIdx.x & 31;
  unsigned got   = 0;
  if (iw == 0)
  {
    constexpr bool compact = sizeof(T) == 4;
    const uint4* src = KIND == 1 ? H.h_in1 : H.h_in2;
    const int n16    = KIND == 1 ? (compact ? nw : 3 * nw) : (nw + 3) / 4;
    long long spins  = 0;
    while (true)
    {
      const unsigned cmd = H.h_cmd[0];
      bool ok            = true;
      for (int e0 = 0; e0 < n16 && ok; e0 += 8 * 32)
      {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          const int e = e0 + 32 * u + lane;
          v[u]        = e < n16 ? ld_sys16(src + e) : make_uint4(0u, 0u, 0u, 0u);
        }
        bool mine = true;
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          const int e = e0 + 32 * u + lane;
          if (e < n16)
          {
            if (KIND == 1)
              mine = mine && (compact ? v[u].w : v[u].z) == seq;
            else
            {
              const unsigned w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (4 * e + q < nw)
                  mine = mine && (w4[q] >> 1) == (seq & 0x7fffffffu);
            }
          }
        }
        ok = __all_sync(0xffffffffu, mine);
        if (ok)
        {
#pragma unroll
          for (int u = 0; u < 8; ++u)
          {
            const int e = e0 + 32 * u + lane;
            if (e < n16)
            {
              if (KIND == 1)
              {
                if constexpr (compact)
                {
                  H.d_displ[3 * e]     = (T)__uint_as_float(v[u].x);
                  H.d_displ[3 * e + 1] = (T)__uint_as_float(v[u].y);
                  H.d_displ[3 * e + 2] = (T)__uint_as_float(v[u].z);
                }
                else
                  H.d_displ[e] = (T)__longlong_as_double((long long)(((unsigned long long)v[u].y << 32) | v[u].x));
              }
              else
              {
                const unsigned w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (4 * e + q < nw)
                    H.d_acc[4 * e + q] = (unsigned char)(w4[q] & 1u);
              }
            }
          }
        }
      }
      if (ok)
      {
        got = seq;
        break;
      }
      if (cmd == SEG_ABORT)
      {
        got = SEG_ABORT;
        break;
      }
      if (++spins > (1ll << 23)) // ~ seconds: the host thread is gone
      {
        got = SEG_ABORT;
        if (lane == 0)
          H.h_cmd[1] = 1u;
        break;
      }
      __nanosleep(100);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0)
      H.d_cmd[0] = got;
  }
  else
  {
    long long spins = 0;
    while (true)
    {
      got = H.d_cmd[0];
      if (got == SEG_ABORT || (int)(got - seq) >= 0)
        break;
      if (++spins > (1ll << 25))
      {
        got = SEG_ABORT;
        break;
      }
      __nanosleep(100);
    }
    __threadfence();
  }
  return got != SEG_ABORT;
}

// grid = live walkers of the crowd (all co-resident); block = SEG_TPB; dynamic smem = seg_layout(...).total
// Moves iat0 .. iat0 + nmoves - 1 of one determinant (rows row0 ..), c0 delays pending at entry, no flush inside.
template<typename T, int CPT, bool HD>
__global__ void __launch_bounds__(SEG_TPB, (sizeof(T) == 4 ? 4 : 2))
    walker_segment_kernel(const __grid_constant__ CUtensorMap tmap, const SplineDev<T> S, const DriverDev<T> Dr,
                          const JastrowDev<T> J, const RngDev R, const SegRng SR, const DetDev<T> D, const int iat0,
                          const int row0, const int nmoves, const int c0, const SegHost<T> H)
{
  // (no static __shared__ variables in this kernel: the ring must sit at a 128-byte aligned shared-memory address for the
  // TMA engine, and the dynamic segment is only guaranteed to start aligned when nothing precedes it)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = D.n, k = D.k, kb = k + 1;
  const int iw = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SegLayout L = seg_layout<T, CPT>(n, k, J.N, J.has_j1 ? J.nions : 0);
  constexpr int STAGE_ELEMS = SEG_ROWS * CPT * SEG_BOXW;
  T* ring   = reinterpret_cast<T*>(smem_raw + L.ring);
  T* vrow   = ring;     // [n]    Ainv[row]      } live only between two gathers, when the ring is idle
  T* glrow  = ring + n; // [3][n] gradient rows  }
  T* phi    = reinterpret_cast<T*>(smem_raw + L.phi); // [5][n] value, gx, gy, gz, lap of the proposed move
  T* x      = reinterpret_cast<T*>(smem_raw + L.x);   // [n] Ainv[row_next] -> inverse row of the move in flight
  T* Bs     = reinterpret_cast<T*>(smem_raw + L.Bs);  // [k][k+1] Woodbury core, resident for the whole segment
  T* pA     = reinterpret_cast<T*>(smem_raw + L.vec); // [k] -V.phi
  T* pB     = pA + k;                                 // [k] U.x
  T* y      = pB + k;
  T* w      = y + k;                                  // [k] w of the prepared row (the accept's bordered update reads it)
  T* red    = reinterpret_cast<T*>(smem_raw + L.red);
  T* rgp    = reinterpret_cast<T*>(smem_raw + L.rgp); // [6][4] per-consumer-warp partial dots
  T* sg     = reinterpret_cast<T*>(smem_raw + L.sg);  // [3] determinant gradient of the prepared row
  T& s_ratio = sg[4];                                 // determinant ratio of the move being decided
  int& s_acc = *reinterpret_cast<int*>(sg + 5);       // its Metropolis decision
  T* s_np    = sg + 8;                                // [3] host-driven mode: the proposed position
  int& s_abort = *reinterpret_cast<int*>(sg + 12);    // host-driven mode: leave the loop
  int mdone  = 0;                                     // moves whose accept / pseudo-accept has been applied
  T* hdr    = reinterpret_cast<T*>(smem_raw + L.hdr); // unit header of the evaluation (spline.cuh)
  uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  uint64_t* empty_bar = full_bar + SEG_NSTAGE;
  uint64_t* v_bar     = empty_bar + SEG_NSTAGE;
  // Between two gathers the ring is idle: the pending rows of V (the stale inverse rows of the delayed electrons) are
  // pulled into it by ONE bulk copy (they are contiguous per walker), and both passes over V -- the accept's -V.phi dots
  // and the next row's x += V^T w -- read shared memory instead of making c dependent trips to L2.
  T* vs                = ring + 4 * n;                                  // behind vrow and glrow
  const int v_cap      = (SEG_NSTAGE * STAGE_ELEMS - 4 * n) / n;         // rows that fit
  const bool v_bulk    = ((size_t)n * sizeof(T)) % 16 == 0;              // bulk copies move multiples of 16 bytes
  unsigned v_phase     = 0;
  T* jred             = reinterpret_cast<T*>(smem_raw + L.jred);
  T* jsum             = reinterpret_cast<T*>(smem_raw + L.jsum);
  unsigned short* jl  = reinterpret_cast<unsigned short*>(smem_raw + L.jl);
  SegMetroPre<T> mpre; // (warp 7)
  // log-determinant increments of the accepted moves: summed here, added to memory once at the end of the segment (a
  // read-modify-write of global memory per accept would put a memory round trip into the accept's critical path)
  double* ld_acc = reinterpret_cast<double*>(smem_raw + L.tim + 192);
  if (tid < 2)
    ld_acc[tid] = 0.0;
#ifdef QMCB_SEG_TIMING
  long long* seg_tim = reinterpret_cast<long long*>(smem_raw + L.tim); // [0,16): cycles per stamp, [16,24): last stamp per warp
  if (threadIdx.x < 24)
    seg_tim[threadIdx.x] = threadIdx.x < 16 ? 0 : clock64();
#endif

  if (tid == 0)
  {
    for (int s = 0; s < SEG_NSTAGE; ++s)
    {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], SEG_NCONS / 32);
    }
    ptx::mbar_init(v_bar, 1);
    ptx::fence_barrier_init();
    if (nmoves > 0)
      ptx::prefetch_tensormap(&tmap);
  }
  if (tid < 3)
    sg[tid] = T(0);
  if (tid == 0)
    s_abort = 0;
  __syncthreads();

  unsigned gq = 0; // ring position; consumers and producer advance it in lock step (SEG_NQ per evaluation)

  for (int m = -1; m < nmoves; ++m)
  {
    const bool part1 = m >= 0, part2 = m + 1 < nmoves;
    const int iat = iat0 + m, row = row0 + m, c = c0 + m; // the move being decided (part1); c = slot it appends
#ifdef QMCB_SEG_TRACE
    const bool trace_on = iw == QMCB_SEG_TRACE && iat0 == 32 && m >= 8 && m < 12;
#endif
    SEG_TR(0); // loop top (after B3 of the previous move)
    const int cn = c + 1;                                   // pending delays when the next row is prepared

    if (part1)
    {
      if constexpr (HD)
      {
        // ---- exchange 1: gradient of the prepared electron out, displacement in (warp 6; everybody else waits at B0)
        if (warp == 6)
        {
          const unsigned seq = H.seq0 + 2u * (unsigned)m;
          const int d        = lane < 3 ? lane : 0;
          T gd               = sg[d];
          if (J.has_j2)
            gd += J.dUat[((size_t)iw * 3 + d) * J.npad + iat];
          if (J.has_j1)
            gd += J.Grad1[((size_t)iw * 3 + d) * J.N + iat];
          const T rold = J.rsoa[((size_t)iw * 3 + d) * J.npad + iat];
          if (lane < 3)
            seg_post(H.h_out1 + 3 * iw + d, gd, seq);
          const bool ok = seg_host_wait<T, 1>(H, iw, Dr.nw, seq);
          if (ok)
          {
            const T p = rold + __ldcg(H.d_displ + 3 * iw + d); // (written by another SM: not through L1)
            if (lane < 3)
            {
              J.newpos[3 * iw + d] = p;
              s_np[d]              = p;
            }
          }
          else if (lane == 0)
            s_abort = 1;
        }
        __syncthreads(); // B0
        // (deferring the Jastrow accept of the previous move into this wait -- warps 0-3 run it while warp 6 waits for the
        // host, the next electron's pair updated first so that its gradient can go out at once -- was measured: e2e 10.0-10.1
        // against 10.5-10.7 M moves/s on the same box, slower)
        if (s_abort)
          break;
      }
      // ======================= gather phase =======================
      if (warp < SEG_NCONS / 32)
      {
        // value, gradient (lattice units) and the Laplacian ALREADY contracted with G G^T: the producer folds the six
        // Hessian prefactors of every (i, j) with the metric into three numbers (A0, A1, A2 below), so a component costs
        // five accumulators instead of ten -- at four CTAs per SM the kernel lives on 64 registers per thread
        T v[CPT], gx[CPT], gy[CPT], gz[CPT], lp[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e)
          v[e] = gx[e] = gy[e] = gz[e] = lp[e] = T(0);
        const int cfirst = tid * CPT;
        const int blk = cfirst / SEG_BOXW, loc = cfirst - blk * SEG_BOXW;
        // Jastrow sums at the proposed position while the first slabs are in flight (every warp derives the proposal
        // itself: a handful of L1/L2 hits and shuffles, no barrier)
        // Jastrow sums at the proposed position while the first slabs are in flight (every warp derives the proposal
        // itself: a handful of L1/L2 hits and shuffles, no barrier).  Interleaving the blocks of the distance pass with the
        // slabs below (one block, one slab, ..., functor pass after the spline epilogue; JastrowMove in jastrow.cuh is
        // resumable for that) was measured: 32.7 vs 30.6 ms of segment kernels per sweep -- the state carried across the
        // slab loop spills at 64 registers per thread
        if (J.has_j2 || J.has_j1)
        {
          T np3[3];
          if constexpr (HD)
            np3[0] = s_np[0], np3[1] = s_np[1], np3[2] = s_np[2];
          else
            seg_propose<T, false>(Dr, J, iw, iat, sg, np3);
          SEG_STAMP(0, 0); // barrier wake-up + proposal
          SEG_TR(5);
          jastrow_move_warps<T>(warp, SEG_JW, J, iw, iat, np3, jl, jred + warp * 16);
          SEG_STAMP(1, 0); // Jastrow sums
          SEG_TR(6);
        }
        T cz[4], dcz[4], d2cz[4];
        // NOT unrolled: the kernel's executed code is ~115 KB (ncu: 7400 SASS instructions touched per launch, "no
        // instruction" the third-largest stall reason) and eight warp roles run different parts of it at once, so it
        // lives or dies by the instruction cache -- the rolled loop is 6 % faster than the 8x unrolled one (26.6 vs 28.4 ms
        // of segment kernels per sweep), and most "harmless" additions to this kernel measured slower for the same reason
#ifdef QMCB_SEG_UNROLLED
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int qq = 0; qq < SEG_NQ; ++qq)
        {
          const unsigned g    = gq + qq;
          const int stage     = g % SEG_NSTAGE;
          const unsigned ph   = (g / SEG_NSTAGE) & 1u;
          ptx::mbar_wait(&full_bar[stage], ph);
          if (qq == 0)
          {
            SEG_TR(7);
            load4<T>(hdr + SPL_HDR_C, cz);
            load4<T>(hdr + SPL_HDR_C + 4, dcz);
            load4<T>(hdr + SPL_HDR_C + 8, d2cz);
          }
          const T* sp = ring + stage * STAGE_ELEMS + blk * (SEG_ROWS * SEG_BOXW) + loc;
          const int i = qq >> 1, j0 = (qq & 1) * 2;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
          {
            T pa[4], pb[4];
            load4<T>(hdr + (i * 4 + j0 + jj) * 8, pa);
            load4<T>(hdr + (i * 4 + j0 + jj) * 8 + 4, pb);
            const T pre00 = pa[0], pre10 = pa[1], pre01 = pa[2], A0 = pa[3], A1 = pb[0], A2 = pb[1];
            const T* p0 = sp + jj * 4 * SEG_BOXW;
            T k0[CPT], k1[CPT], k2[CPT], k3[CPT];
            SegVec<T, CPT>::load(p0, k0);
            SegVec<T, CPT>::load(p0 + SEG_BOXW, k1);
            SegVec<T, CPT>::load(p0 + 2 * SEG_BOXW, k2);
            SegVec<T, CPT>::load(p0 + 3 * SEG_BOXW, k3);
#pragma unroll
            for (int e = 0; e < CPT; ++e)
            {
              const T sum0 = cz[0] * k0[e] + cz[1] * k1[e] + cz[2] * k2[e] + cz[3] * k3[e];
              const T sum1 = dcz[0] * k0[e] + dcz[1] * k1[e] + dcz[2] * k2[e] + dcz[3] * k3[e];
              const T sum2 = d2cz[0] * k0[e] + d2cz[1] * k1[e] + d2cz[2] * k2[e] + d2cz[3] * k3[e];
              lp[e] += A0 * sum0 + A1 * sum1 + A2 * sum2;
              gx[e] += pre10 * sum0;
              gy[e] += pre01 * sum0;
              gz[e] += pre00 * sum1;
              v[e] += pre00 * sum0;
            }
          }
          __syncwarp();
          if (lane == 0)
            ptx::mbar_arrive(&empty_bar[stage]);
        }
        SEG_STAMP(2, 0); // eight slabs
        SEG_TR(8);
        // epilogue: lattice units -> Cartesian, sign, rows into shared memory, dots with the inverse row
        const int bc_sign = *reinterpret_cast<const int*>(hdr + SPL_HDR_SGN);
        const T sgn   = (bc_sign & 1) ? T(-1) : T(1);
        const T dxInv = (T)S.delta_inv[0], dyInv = (T)S.delta_inv[1], dzInv = (T)S.delta_inv[2];
        T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int e = 0; e < CPT; ++e)
        {
          const int mo = cfirst + e;
          if (mo < n)
          {
            const T g0 = gx[e] * dxInv, g1 = gy[e] * dyInv, g2 = gz[e] * dzInv;
            const T psi = sgn * v[e];
            const T dx  = sgn * (S.G[0] * g0 + S.G[1] * g1 + S.G[2] * g2);
            const T dy  = sgn * (S.G[3] * g0 + S.G[4] * g1 + S.G[5] * g2);
            const T dz  = sgn * (S.G[6] * g0 + S.G[7] * g1 + S.G[8] * g2);
            const T lap = sgn * lp[e];
            phi[mo]         = psi;
            phi[n + mo]     = dx;
            phi[2 * n + mo] = dy;
            phi[3 * n + mo] = dz;
            phi[4 * n + mo] = lap;
            const T wr = x[mo];
            acc[0] += psi * wr;
            acc[1] += dx * wr;
            acc[2] += dy * wr;
            acc[3] += dz * wr;
          }
        }
        warp_fold<T, 4>(acc); // lane l holds the total of value l >> 3
        if ((lane & 7) == 0)
          rgp[warp * 4 + (lane >> 3)] = acc[0];
        SEG_TR(9);
      }
      else if (warp == 6)
      {
        // ---- producer: proposal, unit header (spline.cuh producer warp), one TMA request (CPT boxes) per stage
        T np3[3];
        if constexpr (HD)
          np3[0] = s_np[0], np3[1] = s_np[1], np3[2] = s_np[2];
        else
          seg_propose<T, true>(Dr, J, iw, iat, sg, np3);
        SEG_TR(1);
        T* scratch = hdr + SPL_HDR;
        T ru[3];
        const int bc_sign = convert_pos<T, T>(S, np3, ru);
        int my_ind        = 0;
        if (lane < 3)
        {
          const T rud     = lane == 0 ? ru[0] : (lane == 1 ? ru[1] : ru[2]);
          const double di = lane == 0 ? S.delta_inv[0] : (lane == 1 ? S.delta_inv[1] : S.delta_inv[2]);
          const int nmax  = (lane == 0 ? S.M[0] : (lane == 1 ? S.M[1] : S.M[2])) - 1;
          T t, p0[4], p1[4], p2[4];
          get_spline_bound<T>((double)rud * di, nmax, my_ind, t);
          prefactors(p0, p1, p2, t);
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            scratch[lane * 12 + e]     = p0[e];
            scratch[lane * 12 + 4 + e] = p1[e];
            scratch[lane * 12 + 8 + e] = p2[e];
          }
        }
        __syncwarp();
        const int i0 = __shfl_sync(0xffffffffu, my_ind, 0), i1 = __shfl_sync(0xffffffffu, my_ind, 1),
                  i2 = __shfl_sync(0xffffffffu, my_ind, 2);
        if (lane < 16)
        {
          const int i = lane >> 2, j = lane & 3;
          const T ai = scratch[i], dai = scratch[4 + i], d2ai = scratch[8 + i];
          const T bj = scratch[12 + j], dbj = scratch[16 + j], d2bj = scratch[20 + j];
          // metric-weighted Hessian prefactors: lap = sum_ab h_ab GGt_ab with h_ab scaled by the grid spacings
          // (SplineR2R.cpp:349-373; SymTrace, contraction_helper.hpp:47-48)
          const T dxI = (T)S.delta_inv[0], dyI = (T)S.delta_inv[1], dzI = (T)S.delta_inv[2];
          const T wxx = S.symGG[0] * (dxI * dxI), wxy = S.symGG[1] * (dxI * dyI), wxz = S.symGG[2] * (dxI * dzI);
          const T wyy = S.symGG[3] * (dyI * dyI), wyz = S.symGG[4] * (dyI * dzI), wzz = S.symGG[5] * (dzI * dzI);
          T* hd = hdr + lane * 8;
          hd[0] = ai * bj;                                              // pre00: value, d/dz
          hd[1] = dai * bj;                                             // pre10: d/dx, with sum1: d2/dxdz
          hd[2] = ai * dbj;                                             // pre01: d/dy, with sum1: d2/dydz
          hd[3] = (d2ai * bj) * wxx + (dai * dbj) * wxy + (ai * d2bj) * wyy; // A0 (x sum0)
          hd[4] = (dai * bj) * wxz + (ai * dbj) * wyz;                  // A1 (x sum1)
          hd[5] = (ai * bj) * wzz;                                      // A2 (x sum2)
          hd[6] = T(0);
          hd[7] = T(0);
        }
        else if (lane < 19)
        {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            hdr[SPL_HDR_C + (lane - 16) * 4 + e] = scratch[24 + (lane - 16) * 4 + e];
        }
        else if (lane == 19)
          *reinterpret_cast<int*>(hdr + SPL_HDR_SGN) = bc_sign;
        __syncwarp();
        SEG_TR(2);
        const uint64_t pol = ptx::policy_evict_first();
#ifndef QMCB_SEG_SMALL
#define QMCB_SEG_SMALL 1
#endif
#if QMCB_SEG_SMALL
#pragma unroll 1
#endif
        for (int qq = 0; qq < SEG_NQ; ++qq)
        {
          if (qq == 3)
            SEG_TR(3);
          const unsigned g  = gq + qq;
          const int stage   = g % SEG_NSTAGE;
          const unsigned ph = (g / SEG_NSTAGE) & 1u;
          ptx::mbar_wait(&empty_bar[stage], ph ^ 1u);
          if (lane == 0)
          {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], (unsigned)(STAGE_ELEMS * sizeof(T)));
#pragma unroll
            for (int b = 0; b < CPT; ++b)
              ptx::tma_load_4d_hint(ring + stage * STAGE_ELEMS + b * (SEG_ROWS * SEG_BOXW), &tmap, b * SEG_BOXW, i2,
                                    i1 + (qq & 1) * 2, i0 + (qq >> 1), &full_bar[stage], pol);
          }
          __syncwarp();
        }
        SEG_TR(4);
      }
      else
      {
        // ---- warp 7: everything the Metropolis test needs from memory, fetched under the gather; and the rows the NEXT
        // move's preparation reads for the first time since the last flush (the stale inverse row, the gradient rows of
        // that electron, its Gaussians) are pulled into L2 now -- their addresses are known a whole move ahead, and a
        // cold DRAM access inside the boundary chain costs microseconds while the stencil stream keeps HBM busy
#ifdef QMCB_SEG_PREFETCH // (measured: no gain, slightly slower)
        if (part2)
        {
          const char* ar = reinterpret_cast<const char*>(D.Ainv + ((size_t)iw * n + row + 1) * D.lda);
          const char* gr = reinterpret_cast<const char*>(D.GL + ((size_t)iw * n + row + 1) * 4 * n);
          const int la = (int)(((size_t)n * sizeof(T) + 127) / 128), lg = (int)(((size_t)3 * n * sizeof(T) + 127) / 128);
          for (int l = lane; l < la + lg; l += 32)
          {
            const char* pf = l < la ? ar + (size_t)l * 128 : gr + (size_t)(l - la) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
          }
          if (lane == 0)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Dr.deltas + ((size_t)(iat + 1) * Dr.nw + iw) * 3));
        }
#endif
        if constexpr (!HD)
          mpre = seg_metro_prefetch<T>(Dr, J, R, SR, iw, iat, sg);
        if (SEG_JW == 7 && (J.has_j2 || J.has_j1))
        {
          // this warp's share of the Jastrow sums (SEG_JW)
          T np3[3];
          if constexpr (HD)
            np3[0] = s_np[0], np3[1] = s_np[1], np3[2] = s_np[2];
          else
            np3[0] = mpre.np[0], np3[1] = mpre.np[1], np3[2] = mpre.np[2];
          jastrow_move_warps<T>(6, SEG_JW, J, iw, iat, np3, jl, jred + 6 * 16);
        }
        SEG_STAMP(8, 224); // Metropolis prefetch incl. the wait for the previous move's total
        SEG_TR(10);
      }
      gq += SEG_NQ;
      SEG_STAMP(3, 0); // epilogue
      __syncthreads(); // B1
      SEG_TR(11);
      SEG_STAMP(4, 0);   // wait at B1
      SEG_STAMP(9, 224); // wait at B1 (Metropolis warp)
#ifdef QMCB_SEG_TIMING
      if (iw == 194 && iat0 == 32 && m >= 4 && m <= 6 && (tid == 0 || tid == 224 || tid == 100))
        printf("abs iw %d m %d tid %d after B1 clock %lld\n", iw, m, tid, (long long)clock64());
#endif: orbital rows and partial dots in shared memory, Jastrow sums and the proposal in memory
    }

    // ======================= Metropolis test (warp 7) || accept-independent staging and dots (warps 0-6) =======================
    const int cA = part1 ? c : 0;                   // rows of V final before this step
    const int cB = part2 ? (part1 ? c : cn) : 0;    // rows of U final before this step
    const int nv  = part1 ? c : (part2 ? cn : 0);          // rows of V this step reads
    const int nvs = v_bulk ? (nv < v_cap ? nv : v_cap) : 0; // ... of which staged in shared memory
    if (tid == 0 && nvs > 0)
    {
      const unsigned bytes = (unsigned)((size_t)nvs * n * sizeof(T));
      ptx::mbar_arrive_expect_tx(v_bar, bytes);
      ptx::bulk_g2s(vs, D.V + (size_t)iw * k * n, bytes, v_bar);
    }
    if (warp == 7)
    {
      if (part1)
      {
        T q[4] = {T(0), T(0), T(0), T(0)};
        for (int pw = 0; pw < SEG_NCONS / 32; ++pw)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            q[e] += rgp[pw * 4 + e];
        if (J.has_j2 || J.has_j1)
        {
          // the consumer warps' partial Jastrow sums in warp order; stored for the accept (jastrow_accept_body) and API readers
          if (lane < 10)
          {
            T t = T(0);
            for (int pw = 0; pw < SEG_JW; ++pw)
              t += jred[pw * 16 + lane];
            jsum[lane] = t;
            if (lane < 5 && J.has_j2)
              J.j2_vgl[(size_t)iw * 5 + lane] = lane == 4 ? -t : t;
            if (lane >= 5 && J.has_j1)
              J.j1_cur[(size_t)iw * 5 + lane - 5] = t;
          }
          __syncwarp();
        }
        // (forming the Jastrow ratios -- two double-precision exponentials -- on this warp BEFORE the barrier, behind a
        // named-barrier hand-over from the consumers' Jastrow pass, was measured: 29.9 vs 29.5 ms per sweep, slower)
        T rdet;
        bool acc;
        if constexpr (HD)
        {
          // ---- exchange 2: TrialWaveFunction::mw_calcRatioGrad out (ratio as PsiValue = double, gradient of the
          // proposed configuration), accept flag in; the staging and the dot sweeps of warps 0-6 run under the round trip
          const unsigned seq = H.seq0 + 2u * (unsigned)m + 1u;
          rdet               = q[0];
          double ratio       = (double)rdet;
          T gn[3]            = {q[1] / rdet, q[2] / rdet, q[3] / rdet};
          if (J.has_j2)
          {
            ratio = ratio * exp((double)(J.Uat[(size_t)iw * J.npad + iat] - jsum[0]));
            gn[0] += jsum[1], gn[1] += jsum[2], gn[2] += jsum[3];
          }
          if (J.has_j1)
          {
            ratio = ratio * exp((double)(J.Vat[(size_t)iw * J.N + iat] - jsum[5]));
            gn[0] += jsum[6], gn[1] += jsum[7], gn[2] += jsum[8];
          }
          if (lane == 0)
            seg_post(H.h_out2 + 4 * iw, ratio, seq);
          else if (lane < 4)
            seg_post(H.h_out2 + 4 * iw + lane, lane == 1 ? gn[0] : (lane == 2 ? gn[1] : gn[2]), seq);
          const bool ok = seg_host_wait<T, 2>(H, iw, Dr.nw, seq);
          acc           = ok && __ldcg(H.d_acc + iw) != 0;
          if (!ok && lane == 0)
            s_abort = 1;
        }
        else
        {
          SEG_TR(16);
#ifdef QMCB_SEG_TRACE
          acc = seg_metropolis<T>(Dr, J, R, SR, iw, iat, q, jsum, mpre, rdet, trace_on ? &g_seg_trace[m - 8][7][0] : nullptr);
#else
          acc = seg_metropolis<T>(Dr, J, R, SR, iw, iat, q, jsum, mpre, rdet);
#endif
        }
        if (lane == 0)
        {
          s_acc   = acc ? 1 : 0;
          s_ratio = rdet;
        }
        SEG_TR(18);
        SEG_STAMP(10, 224); // Metropolis test
      }
    }
    else
    {
      const Group ga{tid, 7 * 32, 1};
      if (part1)
      {
        const T* arow = D.Ainv + ((size_t)iw * n + row) * D.lda;
        for (int j = tid; j < n; j += ga.n)
          vrow[j] = arow[j];
      }
      if (part2)
      {
        const T* arow = D.Ainv + ((size_t)iw * n + row + 1) * D.lda;
        const T* gl   = D.GL + ((size_t)iw * n + row + 1) * 4 * n;
        for (int j = tid; j < n; j += ga.n)
        {
          x[j]             = arow[j];
          glrow[j]         = gl[j];
          glrow[n + j]     = gl[n + j];
          glrow[2 * n + j] = gl[2 * n + j];
        }
      }
      SEG_TR(12);
      ptx::fence_proxy_async(); // vrow / glrow live in the ring region: generic writes before the next TMA requests
      if (!part1 && c0 > 0)
      {
        // segment entered with pending delays: the core comes from memory once
        const T* B = D.Binv + (size_t)iw * k * k;
        for (int e = tid; e < c0 * k; e += ga.n)
        {
          const int a = e / k, b = e - a * k;
          if (b < c0)
            Bs[a * kb + b] = B[a * k + b];
        }
      }
      ga.sync();
      SEG_TR(13);
      SEG_STAMP(5, 0); // staging of the next row
      const T* Va     = D.V + (size_t)iw * k * n;
      const T* Ub     = D.U + (size_t)iw * k * n;
      const int nspec = (part1 && part2) ? 1 : 0; // phi.x for the slot this move may append
      const int nU = cB + nspec;
      if (v_bulk)
      {
        // rows are multiples of 16 bytes: every lane keeps its slice of x and of phi in registers for the whole task
        // loop and reads two rows at a time with 16-byte loads.  Rows of U (memory) and the speculative dot first, rows
        // of V last -- those come from the staged copy, which lands while the U rows are read.
        constexpr int NV4 = (CPT * SEG_BOXW / 4 + 31) / 32;
        const V4<T> zero4{T(0), T(0), T(0), T(0)};
        V4<T> xr[NV4];
#pragma unroll
        for (int i = 0; i < NV4; ++i)
        {
          const int j4 = 4 * (lane + 32 * i);
          xr[i] = (part2 && j4 < n) ? ld4(x + j4) : zero4;
        }
        auto two_rows = [&](const T* r0, const T* r1, const bool two, const V4<T> (&vec)[NV4], T& s0, T& s1) {
          V4<T> a0[NV4], a1[NV4];
#pragma unroll
          for (int i = 0; i < NV4; ++i)
          {
            const int j4 = 4 * (lane + 32 * i);
            a0[i] = j4 < n ? ld4(r0 + j4) : zero4;
            a1[i] = (two && j4 < n) ? ld4(r1 + j4) : zero4;
          }
          s0 = T(0), s1 = T(0);
#pragma unroll
          for (int i = 0; i < NV4; ++i)
          {
            s0 += dot4(a0[i], vec[i]);
            s1 += dot4(a1[i], vec[i]);
          }
          s0 = warp_sum(s0);
          s1 = warp_sum(s1);
        };
        for (int u0 = warp; u0 < nU; u0 += 14)
        {
          const int u1   = u0 + 7;
          const bool two = u1 < nU;
          const T* r0    = u0 < cB ? Ub + (size_t)u0 * n : phi;
          const T* r1    = two ? (u1 < cB ? Ub + (size_t)u1 * n : phi) : r0;
          T s0, s1;
          two_rows(r0, r1, two, xr, s0, s1);
          if (lane == 0)
          {
            pB[u0] = s0;
            if (two)
              pB[u1] = s1;
          }
        }
        SEG_TR(14);
        if (cA > 0)
        {
          // (the slice of phi replaces the slice of x in the same registers)
#pragma unroll
          for (int i = 0; i < NV4; ++i)
          {
            const int j4 = 4 * (lane + 32 * i);
            xr[i] = j4 < n ? ld4(phi + j4) : zero4;
          }
          if (nvs > 0)
            ptx::mbar_wait(v_bar, v_phase);
          // (the warps take the V rows in reverse order of the U rows so that the work evens out)
          for (int a0i = 6 - warp; a0i < cA; a0i += 14)
          {
            const int a1i  = a0i + 7;
            const bool two = a1i < cA;
            const T* r0    = a0i < nvs ? vs + (size_t)a0i * n : Va + (size_t)a0i * n;
            const T* r1    = two ? (a1i < nvs ? vs + (size_t)a1i * n : Va + (size_t)a1i * n) : r0;
            T s0, s1;
            two_rows(r0, r1, two, xr, s0, s1);
            if (lane == 0)
            {
              pA[a0i] = -s0;
              if (two)
                pA[a1i] = -s1;
            }
          }
        }
      }
      else
      {
        // generic rows (length not a multiple of 16 bytes): scalar loads, everything from memory
        const int ntask = cA + cB + nspec;
        for (int t0 = warp; t0 < ntask; t0 += 14)
        {
          const int t1 = t0 + 7;
          const T *r0, *v0, *r1, *v1;
          auto pick = [&](int t, const T*& r, const T*& v) {
            if (t < cB)
              r = Ub + (size_t)t * n, v = x;
            else if (t < nU)
              r = phi, v = x;
            else
              r = Va + (size_t)(t - nU) * n, v = phi;
          };
          pick(t0, r0, v0);
          r1 = r0, v1 = v0;
          const bool two = t1 < ntask;
          if (two)
            pick(t1, r1, v1);
          T s0(0), s1(0);
          for (int j = lane; j < n; j += 32)
          {
            s0 += r0[j] * v0[j];
            if (two)
              s1 += r1[j] * v1[j];
          }
          s0 = warp_sum(s0);
          s1 = warp_sum(s1);
          if (lane == 0)
          {
            if (t0 < nU)
              pB[t0] = s0;
            else
              pA[t0 - nU] = -s0;
            if (two)
            {
              if (t1 < nU)
                pB[t1] = s1;
              else
                pA[t1 - nU] = -s1;
            }
          }
        }
      }
    }
    SEG_TR(15);
    SEG_STAMP(6, 0); // dots
    __syncthreads(); // B2
    if constexpr (HD)
      if (s_abort)
      {
        if (nvs > 0)
          ptx::mbar_wait(v_bar, v_phase); // (no bulk copy may be in flight into this CTA's shared memory at exit)
        break;
      }
    SEG_TR(19);
    SEG_STAMP(7, 0);    // wait at B2
    SEG_STAMP(11, 224); // wait at B2 (Metropolis warp)
#ifdef QMCB_SEG_TIMING
    if (iw == 194 && iat0 == 32 && m >= 4 && m <= 6 && (tid == 0 || tid == 224 || tid == 100))
      printf("abs iw %d m %d tid %d after B2 clock %lld\n", iw, m, tid, (long long)clock64());
#endif

    // ======================= determinant accept + next row (threads 0 .. SEG_DET-1) || Jastrow accept (the others) =======================
    if (tid < SEG_DET)
    {
      const Group gd{tid, SEG_DET, 1};
      // (pulling the next proposal's cold inputs -- the electron's Gaussians and one-body Jastrow gradient -- into L2 from
      // here was measured: 29.4 vs 28.5 ms of segment kernels per sweep; like every other prefetch tried in this kernel
      // it made things slower)
      if (part1)
      {
        const bool acc = s_acc != 0;
        T* U           = D.U + (size_t)iw * k * n;
        T* Vm          = D.V + (size_t)iw * k * n;
        T* gl          = D.GL + ((size_t)iw * n + row) * 4 * n;
        // rows: V[c] = Ainv[row] for every walker (DelayedUpdateBatched.h:646); U[c] and the G/L rows on accept
        for (int j = tid; j < n; j += gd.n)
        {
          Vm[(size_t)c * n + j] = vrow[j];
          U[(size_t)c * n + j]  = acc ? phi[j] : T(0);
          if (acc)
          {
            st_stream(gl + j, phi[n + j]);
            st_stream(gl + n + j, phi[2 * n + j]);
            st_stream(gl + 2 * n + j, phi[3 * n + j]);
            st_stream(gl + 3 * n + j, phi[4 * n + j]);
          }
        }
        SEG_TR(20);
        if (acc)
        {
          // bordered update of Binv (DelayedUpdate.h:113-141) in shared memory; w is the one left by this row's preparation
          if (tid == 0 && s_ratio == T(0) && D.err)
            atomicOr(D.err, QMCB_ERR_ZERO_RATIO_ACCEPTED);
          const T sigma = T(1) / s_ratio;
          if (tid < c)
          {
            T sacc(0);
            for (int b = 0; b < c; ++b)
              sacc += Bs[tid * kb + b] * pA[b];
            y[tid] = sigma * sacc;
          }
          gd.sync();
          for (int e = tid; e < c * c; e += gd.n)
          {
            const int a = e / c, b = e - a * c;
            Bs[a * kb + b] += y[a] * w[b];
          }
          if (tid < c)
          {
            Bs[tid * kb + c] = y[tid];
            Bs[c * kb + tid] = sigma * w[tid];
          }
          if (tid == 0)
          {
            Bs[c * kb + c]             = sigma;
            D.list[(size_t)iw * k + c] = row;
            logdet_accumulate(ld_acc, s_ratio); // log_value += log(curRatio)
          }
        }
        else
        {
          // pseudo-accept: detail/OMPTarget/AccelMatrixUpdateOMPTarget.hpp:139-160
          if (tid < c)
          {
            Bs[c * kb + tid] = T(0);
            Bs[tid * kb + c] = T(0);
          }
          if (tid == 0)
          {
            Bs[c * kb + c]             = T(1);
            D.list[(size_t)iw * k + c] = -1;
          }
        }
        gd.sync();
      }
      SEG_TR(21);
      if (part2)
      {
        // p'[a] = U[a].x : rows < cB from the staging phase; the appended row is phi.x on accept, 0 for a pseudo-accept
        if (part1 && tid == 0)
          pB[c] = s_acc != 0 ? pB[cB] : T(0);
        gd.sync();
        if (tid < cn)
        {
          T sacc(0);
          for (int a = 0; a < cn; ++a)
            sacc += Bs[a * kb + tid] * pB[a];
          w[tid] = -sacc; // w = -Binv^T p'  (DelayedUpdate.h:100-101)
        }
        gd.sync();
        SEG_TR(22);
        // x += V^T w : the staged rows from shared memory, rows beyond the staging capacity from L2, the row appended
        // by this move from vrow; then the gradient dots with the staged gradient rows
        const T* Vm = D.V + (size_t)iw * k * n;
        T acc3[3]   = {T(0), T(0), T(0)};
        if (nvs > 0)
          ptx::mbar_wait(v_bar, v_phase); // (warps that read no V row in the staging phase have not observed it yet)
        const int cS = cB < nvs ? cB : nvs;
        if (v_bulk)
        {
          // four consecutive columns per thread, 16-byte accesses
#ifndef QMCB_SEG_VTW_UNROLL
#define QMCB_SEG_VTW_UNROLL 1 // (B200, a64: 1 -> 32.7 ms, 4 -> 33.3 ms per sweep: the kernel is register-bound at 64)
#endif
          constexpr int VTW_UNROLL = QMCB_SEG_VTW_UNROLL;
          for (int j4 = 4 * tid; j4 < n; j4 += 4 * gd.n)
          {
            V4<T> sacc{T(0), T(0), T(0), T(0)};
#pragma unroll(VTW_UNROLL)
            for (int a = 0; a < cS; ++a)
            {
              const T wa    = w[a];
              const V4<T> v = ld4(vs + (size_t)a * n + j4);
              sacc.x += v.x * wa, sacc.y += v.y * wa, sacc.z += v.z * wa, sacc.w += v.w * wa;
            }
#pragma unroll 4
            for (int a = cS; a < cB; ++a)
            {
              const T wa    = w[a];
              const V4<T> v = ld4(Vm + (size_t)a * n + j4);
              sacc.x += v.x * wa, sacc.y += v.y * wa, sacc.z += v.z * wa, sacc.w += v.w * wa;
            }
            if (part1)
            {
              const T wc     = w[c];
              const V4<T> vr = ld4(vrow + j4);
              sacc.x += vr.x * wc, sacc.y += vr.y * wc, sacc.z += vr.z * wc, sacc.w += vr.w * wc;
            }
            const V4<T> x0 = ld4(x + j4);
            const V4<T> xv{x0.x + sacc.x, x0.y + sacc.y, x0.z + sacc.z, x0.w + sacc.w};
            x[j4] = xv.x, x[j4 + 1] = xv.y, x[j4 + 2] = xv.z, x[j4 + 3] = xv.w;
            acc3[0] += dot4(xv, ld4(glrow + j4));
            acc3[1] += dot4(xv, ld4(glrow + n + j4));
            acc3[2] += dot4(xv, ld4(glrow + 2 * n + j4));
          }
        }
        else
        {
          for (int j = tid; j < n; j += gd.n)
          {
            T sq(0);
            for (int a = 0; a < cB; ++a)
              sq += Vm[(size_t)a * n + j] * w[a];
            if (part1)
              sq += vrow[j] * w[c];
            const T xv = x[j] + sq;
            x[j]       = xv;
            acc3[0] += xv * glrow[j];
            acc3[1] += xv * glrow[n + j];
            acc3[2] += xv * glrow[2 * n + j];
          }
        }
        SEG_TR(23);
        group_sum<T, 3>(gd, acc3, red);
        if (tid < 3)
          sg[tid] = tid == 0 ? acc3[0] : (tid == 1 ? acc3[1] : acc3[2]);
      }
      // the ring region (vrow, glrow) goes back to the TMA engine
      ptx::fence_proxy_async();
      SEG_TR(24);
      SEG_STAMP(12, 0); // determinant accept + next row
    }
    else if (part1 && s_acc != 0)
    {
#ifdef QMCB_SEG_TRACE
      jastrow_accept_body<T>(Group{tid - SEG_DET, SEG_TPB - SEG_DET, 2}, J, iw, iat, jl,
                             trace_on ? &g_seg_trace[m - 8][SEG_DET / 32][0] : nullptr);
#else
      jastrow_accept_body<T>(Group{tid - SEG_DET, SEG_TPB - SEG_DET, 2}, J, iw, iat, jl);
#endif
    }
    // (pulling the rows the NEXT move's staging reads cold -- stale inverse row, gradient rows, Gaussians of electron
    // row + 2 -- into L2 from this phase was measured too: 1278 vs 1259 us per segment, i.e. slightly slower, like the
    // prefetch under the gather (QMCB_SEG_PREFETCH above); the staging is not DRAM-latency bound)
    SEG_TR(25);
    SEG_STAMP(13, 224); // Jastrow accept
    if (nvs > 0)
      v_phase ^= 1u;
    if (part1)
      ++mdone;
    __syncthreads(); // B3
    SEG_STAMP(14, 0);   // wait at B3 (determinant group)
    SEG_STAMP(15, 224); // wait at B3 (Jastrow group)
#ifdef QMCB_SEG_TIMING
    if (iw == 194 && iat0 == 32 && m >= 4 && m <= 6 && (tid == 0 || tid == 224 || tid == 100))
      printf("abs iw %d m %d tid %d after B3 clock %lld\n", iw, m, tid, (long long)clock64());
#endif
  }

#ifdef QMCB_SEG_TIMING
  if (tid == 0 && (iw % 97) == 0 && iat0 == 32)
    printf("seg_timing iw %d moves %d | propose %lld Jsums %lld slabs %lld epi %lld waitB1 %lld stage %lld dots %lld waitB2 %lld "
           "detacc+prep %lld waitB3 %lld || w7: prefetch %lld waitB1 %lld metro %lld waitB2 %lld Jacc %lld waitB3 %lld\n",
           iw, nmoves, seg_tim[0] / nmoves, seg_tim[1] / nmoves, seg_tim[2] / nmoves, seg_tim[3] / nmoves, seg_tim[4] / nmoves,
           seg_tim[5] / nmoves, seg_tim[6] / nmoves, seg_tim[7] / nmoves, seg_tim[12] / nmoves, seg_tim[14] / nmoves,
           seg_tim[8] / nmoves, seg_tim[9] / nmoves, seg_tim[10] / nmoves, seg_tim[11] / nmoves, seg_tim[13] / nmoves,
           seg_tim[15] / nmoves);
#endif
#ifdef QMCB_SEG_TRACE
  __syncthreads();
  if (iw == QMCB_SEG_TRACE && iat0 == 32 && tid == 0)
    for (int mm = 0; mm < 4; ++mm)
    {
      const long long t0 = g_seg_trace[mm][0][0];
      for (int wv = 0; wv < 8; ++wv)
      {
        printf("trace iw %d move %d warp %d :", iw, 8 + mm, wv);
        for (int id = 0; id < 40; ++id)
          if (g_seg_trace[mm][wv][id] != 0)
            printf(" %d=%lld", id, g_seg_trace[mm][wv][id] - t0);
        printf("\n");
      }
    }
#endif
  // the flush (and any later API call) finds the core and w in memory
  const int cF = c0 + mdone;
  if constexpr (HD)
    if (iw == 0 && tid == 0)
      H.d_cmd[1] = (unsigned)mdone;
  {
    T* B = D.Binv + (size_t)iw * k * k;
    for (int e = tid; e < cF * cF; e += SEG_TPB)
    {
      const int a = e / cF, b = e - a * cF;
      B[a * k + b] = Bs[a * kb + b];
    }
    if (tid < cF)
      D.wvec[(size_t)iw * k + tid] = w[tid];
    if (tid < 2)
      D.logdet[2 * (size_t)iw + tid] += ld_acc[tid];
  }
  // U'[a][j] = sum_b Binv[a][b] V[b][j] for the flush that follows (DelayedUpdateBatched.h:716-722, the second of the
  // three products): the core is in shared memory here and the rows of V are L2-hot, so the flush kernel finds U' ready
  // in memory instead of every one of its CTAs re-deriving it.  Thread -> (column j, sixteen slots a).
  {
    const T* Vg = D.V + (size_t)iw * k * n;
    T* Upg      = D.Up + (size_t)iw * k * n;
    const int nh = (cF + 15) / 16;
    for (int item = tid; item < n * nh; item += SEG_TPB)
    {
      const int h = item / n, j = item - h * n, a0 = 16 * h;
      T acc[16];
#pragma unroll
      for (int a = 0; a < 16; ++a)
        acc[a] = T(0);
      for (int b = 0; b < cF; ++b)
      {
        const T v = Vg[(size_t)b * n + j];
#pragma unroll
        for (int a = 0; a < 16; ++a)
          if (a0 + a < cF)
            acc[a] += Bs[(a0 + a) * kb + b] * v;
      }
#pragma unroll
      for (int a = 0; a < 16; ++a)
        if (a0 + a < cF)
          Upg[(size_t)(a0 + a) * n + j] = acc[a];
    }
  }
}
#endif // __CUDACC__

} // namespace qmcb
