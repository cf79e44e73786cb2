// qmcpack_b200/csrc/crowd.cu -- the walker batch ("crowd"): per-walker wavefunction state in HBM and the multi-walker
// operations of the reference's TrialWaveFunction / DiracDeterminantBatched / TwoBodyJastrow / ParticleSet on it.
#include "internal.h"
#include "spline.cuh"
#include "det.cuh"
#include "jastrow.cuh"
#include "driver.cuh"
#include "segment.cuh"
#include "woodbury.cuh"
#include "woodbury_tc5.cuh"
#include "woodbury_dmma.cuh"
#include "inverse.cuh"
#include <cstdlib>
#include <cuda.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <type_traits>
#include <cublas_v2.h>
#include <nvtx3/nvToolsExt.h>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>
#include <map>
#include <mutex>

namespace qmcb
{
// NVTX ranges named after the reference's timers (Utilities/NewTimer.cpp:48-49,94-95 pushes an NVTX range per ScopedTimer;
// names: DiracDeterminantBase.h:42-47, TrialWaveFunction.cpp TWF_timers_): visible in Nsight Systems, free otherwise
struct NvtxRange
{
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define QMCB_NVTX(name) qmcb::NvtxRange nvtx_range__(name)

#define QMCB_CUBLAS(call)                                                                      \
  do                                                                                           \
  {                                                                                            \
    cublasStatus_t s__ = (call);                                                               \
    if (s__ != CUBLAS_STATUS_SUCCESS)                                                          \
      throw std::runtime_error(std::string(#call) + " failed with cuBLAS status " + std::to_string((int)s__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------------------
// small glue kernels of the trial wavefunction (component sums)
// ------------------------------------------------------------------------------------------------------------
// TrialWaveFunction::mw_evalGrad: grads_now = det + J2 dUat[iat] + J1 Grad[iat]  (component order det, J2, J1)
// T = real type of positions and Jastrows, V = determinant value type (T, or cx<T> with complex orbitals)
template<typename T, typename V>
__global__ void twf_grad_kernel(const JastrowDev<T> J, const int iat, const V* det_grads, V* grads_now)
{
  const int iw = blockIdx.x * blockDim.x + threadIdx.x;
  if (iw >= J.nw)
    return;
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    V g = det_grads ? det_grads[3 * iw + d] : V(0);
    if (J.has_j2)
      g += V(J.dUat[((size_t)iw * 3 + d) * J.npad + iat]);
    if (J.has_j1)
      g += V(J.Grad1[((size_t)iw * 3 + d) * J.N + iat]);
    grads_now[3 * iw + d] = g;
  }
}

// ParticleSet::mw_makeMove: newpos = R[iat] + displ
template<typename T>
__global__ void make_move_kernel(const JastrowDev<T> J, const int iat, const T* displ)
{
  const int iw = blockIdx.x * blockDim.x + threadIdx.x;
  if (iw >= J.nw)
    return;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    J.newpos[3 * iw + d] = J.rsoa[((size_t)iw * 3 + d) * J.npad + iat] + displ[3 * iw + d];
}

// host-driven path: ParticleSet::mw_makeMove and the proposed move's Jastrow sums in ONE launch on the crowd stream (one
// CTA per walker).  The alternative -- make_move_kernel, then jastrow_move_kernel on a side stream beside the spline
// gather -- costs one more launch and four event operations per move, and the driver calls, not the GPU, bound that path.
template<typename T>
__global__ void __launch_bounds__(JAS_TPB) make_move_jastrow_kernel(const JastrowDev<T> J, const int iat, const T* displ)
{
  __shared__ T red[10 * 32];
  __shared__ unsigned short jl[JAS_LIST];
  __shared__ T s_pos[3];
  const int iw = blockIdx.x;
  if (threadIdx.x < 3)
  {
    const T p = J.rsoa[((size_t)iw * 3 + threadIdx.x) * J.npad + iat] + displ[3 * iw + threadIdx.x];
    J.newpos[3 * iw + threadIdx.x] = p;
    s_pos[threadIdx.x]             = p;
  }
  __syncthreads();
  const T pos[3] = {s_pos[0], s_pos[1], s_pos[2]};
  jastrow_move_body<T, false>(J, iw, iat, pos, red, jl);
}

// TrialWaveFunction::mw_calcRatioGrad combination for one walker
// (PsiValue ratio = double or cx<double>; the Jastrow factors are real)
template<typename T, typename V>
__device__ __forceinline__ typename value_traits<V>::dbl_t twf_ratio_grad(const JastrowDev<T>& J, int iw, int iat, const V* rg,
                                                                          int nparts, V gn[3])
{
  using PsiV = typename value_traits<V>::dbl_t;
  V q[4];
  sum_rg_parts<V, 4>(rg, iw, nparts, q);
  const V rdet = q[0];
  PsiV ratio   = to_dbl(rdet);
  gn[0]        = q[1] / rdet;
  gn[1]        = q[2] / rdet;
  gn[2]        = q[3] / rdet;
  if (J.has_j2)
  {
    const T* vgl = J.j2_vgl + (size_t)iw * 5;
    ratio        = ratio * exp((double)(J.Uat[(size_t)iw * J.npad + iat] - vgl[0]));
    gn[0] += V(vgl[1]);
    gn[1] += V(vgl[2]);
    gn[2] += V(vgl[3]);
  }
  if (J.has_j1)
  {
    const T* cur = J.j1_cur + (size_t)iw * 5;
    ratio        = ratio * exp((double)(J.Vat[(size_t)iw * J.N + iat] - cur[0]));
    gn[0] += V(cur[1]);
    gn[1] += V(cur[2]);
    gn[2] += V(cur[3]);
  }
  return ratio;
}

template<typename T, typename V>
__global__ void twf_ratio_kernel(const JastrowDev<T> J, const int iat, const V* rg, const int nparts,
                                 typename value_traits<V>::dbl_t* ratios, V* grads)
{
  const int iw = blockIdx.x * blockDim.x + threadIdx.x;
  if (iw >= J.nw)
    return;
  V gn[3];
  ratios[iw]       = twf_ratio_grad<T, V>(J, iw, iat, rg, nparts, gn);
  grads[3 * iw]     = gn[0];
  grads[3 * iw + 1] = gn[1];
  grads[3 * iw + 2] = gn[2];
}

// J2-only ratio/grad for the component-level API
template<typename T>
__global__ void j2_ratio_kernel(const JastrowDev<T> J, const int iat, double* ratios, T* grads)
{
  const int iw = blockIdx.x * blockDim.x + threadIdx.x;
  if (iw >= J.nw)
    return;
  const T* vgl = J.j2_vgl + (size_t)iw * 5;
  ratios[iw]   = exp((double)(J.Uat[(size_t)iw * J.npad + iat] - vgl[0]));
  grads[3 * iw]     = vgl[1];
  grads[3 * iw + 1] = vgl[2];
  grads[3 * iw + 2] = vgl[3];
}

// ------------------------------------------------------------------------------------------------------------
// device driver: everything that happens for ONE walker between two spline evaluations, in one CTA:
//   part 1 (electron iat_prev):  Metropolis test (VMCBatched.cpp:139-167) + determinant accept / pseudo-accept
//                                + Jastrow accept + position commit
//   part 2 (electron iat_next):  inverse-row preparation + old gradient + UNR drift + proposed position
// The only cross-walker dependency of the reference's loop is the ORDER in which walkers draw their uniform from the
// crowd's generator (drawn only when prob >= eps, walkers in index order).  Each CTA publishes "needs a draw" in a
// flag tagged with the move's epoch and looks back over the lower-index walkers (which the hardware dispatched earlier)
// to find its position in the stream: no grid-wide barrier, no single-CTA kernel.
// 256 threads: threads 0-127 determinant group, 128-255 Jastrow group; dynamic smem (n + 2k) * sizeof(T).
// ------------------------------------------------------------------------------------------------------------
constexpr int MB_TPB = 256;

// Metropolis test of walker iw for electron iat_prev by ONE warp (all 32 lanes call it); returns acc and the determinant
// ratio in lane 0
__device__ __forceinline__ float shfl_value(const float v, const int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_value(const double v, const int src) { return __shfl_sync(0xffffffffu, v, src); }
template<typename T>
__device__ __forceinline__ cx<T> shfl_value(const cx<T>& v, const int src)
{
  return cx<T>(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}

// SFNBranch::phaseChanged(std::arg(ratio)) || ratio == 0  (DMCBatched.cpp:188-193, SFNBranch.h:161-169)
__device__ __forceinline__ bool phase_rejects(const double ratio) { return !(ratio > 0.0); }
__device__ __forceinline__ bool phase_rejects(const cx<double>& ratio) { return ratio.re == 0.0 && ratio.im == 0.0; }

template<typename T, typename V>
__device__ __forceinline__ bool metropolis_warp(const DriverDev<T>& Dr, const JastrowDev<T>& J, const RngDev& R, const int iw,
                                                const int iat_prev, const V* rg, const int rg_nparts, V& rdet_out)
{
  const int lane = threadIdx.x & 31;
  // partial ratio/gradient dots of the spline kernel: lanes fetch, every lane adds in index order
  V q[4] = {V(0), V(0), V(0), V(0)};
  {
    V mine[4] = {V(0), V(0), V(0), V(0)};
    if (lane < rg_nparts)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        mine[e] = rg[((size_t)iw * rg_nparts + lane) * 4 + e];
    for (int part = 0; part < rg_nparts && part < 32; ++part)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        q[e] += shfl_value(mine[e], part);
    for (int part = 32; part < rg_nparts; ++part) // more than 32 parts (very wide determinants)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        q[e] += rg[((size_t)iw * rg_nparts + part) * 4 + e];
  }
  const unsigned long long base = R.pos[iat_prev & 1];
  const unsigned sweep          = *R.sweep;
  const V rdet = q[0];
  // PsiValue ratio (double / complex double) and the REAL part of the new gradient: a complex gradient reaches the drift
  // through convertToReal (DriftModifierUNR.cpp:20-23)
  typename value_traits<V>::dbl_t ratio = to_dbl(rdet);
  T gn[3] = {real_part(q[1] / rdet), real_part(q[2] / rdet), real_part(q[3] / rdet)};
  if (J.has_j2)
  {
    const T* vgl = J.j2_vgl + (size_t)iw * 5;
    ratio        = ratio * exp((double)(J.Uat[(size_t)iw * J.npad + iat_prev] - vgl[0]));
    gn[0] += vgl[1];
    gn[1] += vgl[2];
    gn[2] += vgl[3];
  }
  if (J.has_j1)
  {
    const T* cur = J.j1_cur + (size_t)iw * 5;
    ratio        = ratio * exp((double)(J.Vat[(size_t)iw * J.N + iat_prev] - cur[0]));
    gn[0] += cur[1];
    gn[1] += cur[2];
    gn[2] += cur[3];
  }
  T log_gf = T(0), log_gb = T(0);
  if (Dr.use_drift)
  {
    const T* dl = Dr.delta_cur + 3 * iw;
    log_gf      = -Dr.oneover2tau * (dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    T dr[3];
    get_drift<T>(Dr.tauovermass, gn, dr);
    dr[0] += Dr.drifts[3 * iw];
    dr[1] += Dr.drifts[3 * iw + 1];
    dr[2] += Dr.drifts[3 * iw + 2];
    log_gb = -Dr.oneover2tau * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
  }
  const T eps = sizeof(T) == 4 ? T(1.1920929e-07f) : T(2.220446049250313e-16);
  T prob      = (T)norm2(ratio); // std::norm(ratio), VMCBatched.cpp:152
  if (isnan(norm2(ratio)) && lane == 0 && Dr.err)
    atomicOr(Dr.err, QMCB_ERR_NAN_RATIO); // NaNguard::checkOneParticleRatio (TrialWaveFunction.cpp:549): the host throws
  bool need   = prob >= eps;     // periodic cell: every move is valid
  T rr        = T(0);
  if (Dr.dmc)
  {
    // DMCBatched.cpp:188-250: rr = tau |delta|^2 of the raw Gaussian; reject on a zero ratio or a phase change
    // (SFNBranch::phaseChanged: cos(arg ratio) < eps for real wavefunctions = node crossing; never for complex ones);
    // prob = |ratio|^2 exp(log_gb - log_gf) as a whole against eps and the uniform
    const T* dr = Dr.deltas + ((size_t)iat_prev * Dr.nw + iw) * 3;
    rr          = Dr.tauovermass * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
    const bool reject = phase_rejects(ratio);
    prob              = (T)(norm2(ratio) * (double)exp(log_gb - log_gf));
    need              = !reject && prob >= eps;
  }
  // ---- position of this walker's draw in the crowd's stream: look back over the lower-index walkers
  const unsigned epoch = (sweep * (unsigned)(Dr.N + 1) + (unsigned)iat_prev + 1u) & 0x7fffffffu; // 31 bits: the flag word keeps bit 0 for `need`
  if (lane == 0)
    *((volatile unsigned*)(R.flags + iw)) = (epoch << 1) | (need ? 1u : 0u);
  unsigned cnt = 0;
  for (int j = lane; j < iw; j += 32)
  {
    unsigned f;
    do
    {
      f = *((volatile unsigned*)(R.flags + j));
    } while ((f >> 1) != epoch);
    cnt += f & 1u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  bool acc = false;
  if (lane == 0)
  {
    if (need)
    {
      const double u = rng_uniform(R, base + cnt);
      acc            = Dr.dmc ? (u < (double)prob) : (u < (double)(prob * exp(log_gb - log_gf)));
    }
    if (iw == Dr.nw - 1)
      R.pos[(iat_prev + 1) & 1] = base + cnt + (need ? 1u : 0u);
    Dr.accepted[iw] = acc ? 1 : 0;
    if (Dr.dmc)
    {
      Dr.rr_proposed[iw] += rr;
      if (acc)
        Dr.rr_accepted[iw] += rr;
    }
    if (acc)
      Dr.n_accept[iw] += 1;
    else
      Dr.n_reject[iw] += 1;
    if (Dr.accept_log)
      Dr.accept_log[(size_t)iat_prev * Dr.nw + iw] = acc ? 1 : 0;
  }
  rdet_out = rdet;
  return acc;
}

// Phase A (independent of the Metropolis decision, overlapped with it):
//   warp 7: Metropolis test.   warps 0-6 stage into shared memory: phi_vgl rows of the proposed move, x = Ainv[row_next],
//   vrow = Ainv[row_prev] (the future V[c]), the gradient rows of electron row_next, Binv and the previous w; then run
//   the big dot sweeps pA[a] = -V[a].phi (accept) and pB[a] = U[a].x (next row preparation) over the rows that are
//   already final, two rows per warp iteration for memory-level parallelism, plus phi.x for the slot being appended.
// Phase B: threads 0-127: bordered update of Binv entirely in shared memory, row stores (U[c], V[c], G/L rows), then
//   w = -Binv^T p, x += V^T w (one pass over V), inverse-row store and the gradient dot from the staged rows;
//   threads 128-255: Jastrow accept + position commit.   Then warp 0 proposes the next move.
template<typename T, typename V>
__global__ void __launch_bounds__(MB_TPB, (sizeof(V) > 8 ? 2 : 4))
    move_boundary_kernel(const DriverDev<T> Dr, const JastrowDev<T> J, const RngDev R, const DetDev<V> Dacc,
                         const int iat_prev, const int row_prev, const int c_prev, const V* rg, const int rg_nparts,
                         const V* phi_vgl, const DetDev<V> Dprep, const int iat_next, const int row_next, const int c_next,
                         V* det_grads_out, const unsigned char* ext_accept, V* twf_grads_out)
{
  // ext_accept != nullptr: host-driven mode (the Metropolis test ran on the host, flags come from the caller) and
  // twf_grads_out receives the component-summed old gradient instead of a device-side proposal
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ V red[3 * 32];
  __shared__ int s_acc;
  __shared__ V s_ratio;
  __shared__ T jred[10 * 32];
  __shared__ unsigned short jlist[JAS_LIST]; // cutoff lists of the Jastrow passes (jastrow.cuh)
  __shared__ T s_newpos[3];
  __shared__ int s_iw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int iw = blockIdx.x;
  if (Dr.ticket)
  {
    // walker index = start order of the CTA (driver.cuh: DriverDev::ticket); the CTA that draws the last ticket of the
    // launch re-arms the counter for the next one (launches of a crowd are stream ordered)
    if (tid == 0)
    {
      const unsigned t = atomicAdd(Dr.ticket, 1u);
      if (t == gridDim.x - 1)
        *((volatile unsigned*)Dr.ticket) = 0u;
      s_iw = (int)t;
    }
    __syncthreads();
    iw = s_iw;
  }
  const bool part1 = iat_prev >= 0, part2 = iat_next >= 0;
  const int nA = part1 ? Dacc.n : 0, nB = part2 ? Dprep.n : 0, k = part1 ? Dacc.k : Dprep.k;
  const int kb = k + 1; // padded row stride of the staged Binv (conflict-free row AND column walks)
  V* phi   = reinterpret_cast<V*>(smem_raw); // [5][nA]  value, gx, gy, gz, lap of the proposed move
  V* vrow  = phi + 5 * nA;                   // [nA]     Ainv[row_prev]
  V* x     = vrow + nA;                      // [nB]     Ainv[row_next] -> inverse row
  V* glrow = x + nB;                         // [3][nB]  gradient rows of electron row_next
  V* Bs    = glrow + 3 * nB;                 // [k][k+1] Binv of the determinant being updated / prepared
  V* pA    = Bs + k * kb;                    // [k]  accept:  -V.phi
  V* pB    = pA + k;                         // [k]  prepare:  U.x
  V* y     = pB + k;                         // [k]
  V* w     = y + k;                          // [k]  new w
  V* wold  = w + k;                          // [k]  w left by the previous preparation (needed by the bordered update)
  // a spin change or a flush always splits the boundary, so when both parts are present they share the determinant
  const bool same_det = part1 && part2;
  const int cA = part1 ? c_prev : 0;                          // rows of V for the accept dots
  const int cB = part2 ? (same_det ? c_prev : c_next) : 0;    // rows of U that are final before this kernel
  const DetDev<V>& Dm = part1 ? Dacc : Dprep;                 // determinant whose Binv is staged
  const int cM        = part1 ? c_prev : c_next;              // its pending count before this kernel
  if (Dr.pdl_early)
    pdl_launch_dependents();
  pdl_wait(); // (no-op unless launched as a programmatic dependent of the spline gather)

  // The Jastrow passes of phase B and of the tail start chains of dependent global loads (positions -> per-particle
  // sums -> functor tables) long after the kernel began; their inputs are known now, so the lines are pulled into L1
  // while phase A runs (one prefetch per 128-byte line, no registers held)
  if (Dr.l1_prefetch && (J.has_j2 || J.has_j1))
  {
    const int lines_row = (J.npad * (int)sizeof(T) + 127) / 128;
    const char* rs = reinterpret_cast<const char*>(J.rsoa + (size_t)iw * 3 * J.npad);
    for (int l = tid; l < 3 * lines_row; l += MB_TPB)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(rs + (size_t)l * 128));
    if (part1 && J.has_j2)
    {
      const char* ua = reinterpret_cast<const char*>(J.Uat + (size_t)iw * J.npad);
      const char* du = reinterpret_cast<const char*>(J.dUat + (size_t)iw * 3 * J.npad);
      const char* d2 = reinterpret_cast<const char*>(J.d2Uat + (size_t)iw * J.npad);
      for (int l = tid; l < 5 * lines_row; l += MB_TPB)
      {
        const char* p = l < lines_row ? ua + (size_t)l * 128
                                      : (l < 4 * lines_row ? du + (size_t)(l - lines_row) * 128 : d2 + (size_t)(l - 4 * lines_row) * 128);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
      }
    }
  }

  if (warp == 7)
  {
    if (part1 && ext_accept)
    {
      if (lane == 0)
      {
        V q[4];
        sum_rg_parts<V, 4>(rg, iw, rg_nparts, q);
        s_acc   = ext_accept[iw] ? 1 : 0;
        s_ratio = q[0];
      }
    }
    else if (part1)
    {
      V rdet;
      const bool acc = metropolis_warp<T, V>(Dr, J, R, iw, iat_prev, rg, rg_nparts, rdet);
      if (lane == 0)
      {
        s_acc   = acc ? 1 : 0;
        s_ratio = rdet;
      }
    }
  }
  else
  {
    const Group ga{tid, 7 * 32, 1};
    if (part1)
    {
      const size_t fs = (size_t)Dacc.nw * nA;
      const V* ph     = phi_vgl + (size_t)iw * nA;
      const V* arow   = Dacc.Ainv + ((size_t)iw * nA + row_prev) * Dacc.lda;
      for (int j = tid; j < nA; j += ga.n)
      {
        phi[j]          = ld_stream(ph + j); // read once: streaming loads (common.cuh)
        phi[nA + j]     = ld_stream(ph + fs + j);
        phi[2 * nA + j] = ld_stream(ph + 2 * fs + j);
        phi[3 * nA + j] = ld_stream(ph + 3 * fs + j);
        phi[4 * nA + j] = ld_stream(ph + 4 * fs + j);
        vrow[j]         = arow[j];
      }
      const V* wv = Dacc.wvec + (size_t)iw * k;
      for (int b = tid; b < cA; b += ga.n)
        wold[b] = wv[b];
    }
    if (part2)
    {
      const V* arow = Dprep.Ainv + ((size_t)iw * nB + row_next) * Dprep.lda;
      const V* gl   = Dprep.GL + ((size_t)iw * nB + row_next) * 4 * nB;
      for (int j = tid; j < nB; j += ga.n)
      {
        x[j]              = arow[j];
        glrow[j]          = gl[j];
        glrow[nB + j]     = gl[nB + j];
        glrow[2 * nB + j] = gl[2 * nB + j];
      }
    }
    {
      const V* B = Dm.Binv + (size_t)iw * k * k;
      for (int e = tid; e < cM * k; e += ga.n)
      {
        const int a = e / k, b = e - a * k;
        if (b < cM)
          Bs[a * kb + b] = B[a * k + b];
      }
    }
    ga.sync();
    const V* Va = part1 ? Dacc.V + (size_t)iw * k * nA : nullptr;
    const V* Ub = part2 ? Dprep.U + (size_t)iw * k * nB : nullptr;
    const int nspec = same_det ? 1 : 0; // phi.x for the slot this kernel appends (shared memory only)
    const int ntask = cA + cB + nspec;
    for (int t0 = warp; t0 < ntask; t0 += 14)
    {
      const int t1 = t0 + 7;
      const V *r0, *v0, *r1, *v1;
      int n0, n1 = 0;
      auto pick = [&](int t, const V*& r, const V*& v, int& nn) {
        if (t < cA)
          r = Va + (size_t)t * nA, v = phi, nn = nA;
        else if (t < cA + cB)
          r = Ub + (size_t)(t - cA) * nB, v = x, nn = nB;
        else
          r = phi, v = x, nn = nB;
      };
      pick(t0, r0, v0, n0);
      r1 = r0, v1 = v0;
      if (t1 < ntask)
        pick(t1, r1, v1, n1);
      V s0(0), s1(0);
      const int nmax = n0 > n1 ? n0 : n1;
      // DOT_B elements of both rows are requested before the first multiply: 2 * DOT_B loads in flight per lane, one
      // memory round trip per 32 * DOT_B columns instead of one per unrolled group (same accumulation order)
      constexpr int DOT_B = sizeof(V) > 8 ? 4 : (sizeof(V) > 4 ? 6 : 12);
      for (int jb = lane; jb < nmax; jb += 32 * DOT_B)
      {
        V a0[DOT_B], a1[DOT_B];
#pragma unroll
        for (int q = 0; q < DOT_B; ++q)
        {
          const int j = jb + 32 * q;
          a0[q]       = j < n0 ? r0[j] : V(0);
          a1[q]       = j < n1 ? r1[j] : V(0);
        }
#pragma unroll
        for (int q = 0; q < DOT_B; ++q)
        {
          const int j = jb + 32 * q;
          if (j < n0)
            s0 += a0[q] * v0[j];
          if (j < n1)
            s1 += a1[q] * v1[j];
        }
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0)
      {
        if (t0 < cA)
          pA[t0] = -s0;
        else
          pB[t0 - cA] = s0;
        if (t1 < ntask)
        {
          if (t1 < cA)
            pA[t1] = -s1;
          else
            pB[t1 - cA] = s1;
        }
      }
    }
  }
  __syncthreads();

  V g[3] = {V(0), V(0), V(0)};
  if (tid < MB_TPB / 2)
  {
    const Group gd{tid, MB_TPB / 2, 1};
    int cN = cM; // pending rows of the staged determinant after part 1
    if (part1)
    {
      const bool acc = s_acc != 0;
      const int c    = c_prev;
      V* U           = Dacc.U + (size_t)iw * k * nA;
      V* Vm          = Dacc.V + (size_t)iw * k * nA;
      // rows: V[c] = Ainv[row_prev] for every walker (DelayedUpdateBatched.h:646); U[c] and the G/L rows on accept
      V* gl = Dacc.GL + ((size_t)iw * nA + row_prev) * 4 * nA;
      for (int j = tid; j < nA; j += gd.n)
      {
        Vm[(size_t)c * nA + j] = vrow[j];
        U[(size_t)c * nA + j]  = acc ? phi[j] : V(0);
        if (acc && !Dr.value_only)
        {
          st_stream(gl + j, phi[nA + j]); // next read a whole sweep later: streaming stores
          st_stream(gl + nA + j, phi[2 * nA + j]);
          st_stream(gl + 2 * nA + j, phi[3 * nA + j]);
          st_stream(gl + 3 * nA + j, phi[4 * nA + j]);
        }
      }
      if (acc)
      {
        // bordered update of Binv (DelayedUpdate.h:113-141) on the staged copy
        if (tid == 0 && s_ratio == V(0) && Dacc.err)
          atomicOr(Dacc.err, QMCB_ERR_ZERO_RATIO_ACCEPTED);
        const V sigma = V(1) / s_ratio;
        if (tid < c)
        {
          V sacc(0);
          for (int b = 0; b < c; ++b)
            sacc += Bs[tid * kb + b] * pA[b];
          y[tid] = sigma * sacc;
        }
        gd.sync();
        for (int e = tid; e < c * c; e += gd.n)
        {
          const int a = e / c, b = e - a * c;
          Bs[a * kb + b] += y[a] * wold[b];
        }
        if (tid < c)
        {
          Bs[tid * kb + c] = y[tid];
          Bs[c * kb + tid] = sigma * wold[tid];
        }
        if (tid == 0)
        {
          Bs[c * kb + c]                = sigma;
          Dacc.list[(size_t)iw * k + c] = row_prev;
          logdet_accumulate(Dacc.logdet + 2 * (size_t)iw, s_ratio); // log_value += log(curRatio)
        }
      }
      else
      {
        // pseudo-accept: detail/OMPTarget/AccelMatrixUpdateOMPTarget.hpp:139-160
        if (tid < c)
        {
          Bs[c * kb + tid] = V(0);
          Bs[tid * kb + c] = V(0);
        }
        if (tid == 0)
        {
          Bs[c * kb + c]                = V(1);
          Dacc.list[(size_t)iw * k + c] = -1;
        }
      }
      cN = c + 1;
      gd.sync();
      // write the updated core back (rows/columns < cN)
      V* B = Dacc.Binv + (size_t)iw * k * k;
      for (int e = tid; e < cN * k; e += gd.n)
      {
        const int a = e / k, b = e - a * k;
        if (b < cN)
          B[a * k + b] = Bs[a * kb + b];
      }
    }
    if (part2)
    {
      // p'[a] = U[a].x : rows < cB from phase A; the appended row is phi.x on accept and 0 for a pseudo-accept
      if (same_det && tid == 0)
        pB[c_prev] = s_acc != 0 ? pB[cB] : V(0);
      gd.sync();
      // w = -Binv^T p'  (DelayedUpdate.h:100-101), kept for the accept of this electron
      if (tid < cN)
      {
        V sacc(0);
        for (int a = 0; a < cN; ++a)
          sacc += Bs[a * kb + tid] * pB[a];
        w[tid]                            = -sacc;
        Dprep.wvec[(size_t)iw * k + tid] = -sacc;
      }
      gd.sync();
      // x += V^T w : rows < cB from HBM/L2, the row appended by part 1 from shared memory
      const V* Vm = Dprep.V + (size_t)iw * k * nB;
      V acc3[3]   = {V(0), V(0), V(0)};
      V* out      = Dprep.invRow + (size_t)iw * nB;
      // XC columns per thread side by side: the a-loop then carries XC independent load streams (the stores to out[]
      // would otherwise keep the compiler from overlapping the columns); same per-column accumulation order
      constexpr int XC = sizeof(V) > 8 ? 2 : 3;
      for (int jb = tid; jb < nB; jb += XC * gd.n)
      {
        V sacc[XC];
#pragma unroll
        for (int q = 0; q < XC; ++q)
          sacc[q] = V(0);
#pragma unroll 4
        for (int a = 0; a < cB; ++a)
        {
          const V wa = w[a];
#pragma unroll
          for (int q = 0; q < XC; ++q)
          {
            const int j = jb + q * gd.n;
            if (j < nB)
              sacc[q] += Vm[(size_t)a * nB + j] * wa;
          }
        }
#pragma unroll
        for (int q = 0; q < XC; ++q)
        {
          const int j = jb + q * gd.n;
          if (j < nB)
          {
            V sq = sacc[q];
            if (same_det)
              sq += vrow[j] * w[c_prev];
            const V xv = x[j] + sq;
            out[j]     = xv;
            acc3[0] += xv * glrow[j];
            acc3[1] += xv * glrow[nB + j];
            acc3[2] += xv * glrow[2 * nB + j];
          }
        }
      }
      if (Dr.use_drift)
      {
        group_sum<V, 3>(gd, acc3, red);
        g[0] = acc3[0];
        g[1] = acc3[1];
        g[2] = acc3[2];
      }
    }
  }
  else if (part1 && s_acc != 0)
    jastrow_accept_body<T>(Group{tid - MB_TPB / 2, MB_TPB / 2, 2}, J, iw, iat_prev, jlist);
  __syncthreads();

  if (part2 && twf_grads_out)
  {
    // host-driven mode: TrialWaveFunction::mw_evalGrad = determinant + J2 + J1 gradients of electron iat_next
    if (tid < 3)
    {
      V gd = tid == 0 ? g[0] : (tid == 1 ? g[1] : g[2]);
      if (J.has_j2)
        gd += V(J.dUat[((size_t)iw * 3 + tid) * J.npad + iat_next]);
      if (J.has_j1)
        gd += V(J.Grad1[((size_t)iw * 3 + tid) * J.N + iat_next]);
      twf_grads_out[3 * iw + tid] = gd;
    }
  }
  else if (part2 && tid < 32)
  {
    // proposal: lanes 0-2 own one Cartesian component each (independent loads), |g|^2 through shuffles
    const int d = lane < 3 ? lane : 0;
    const T delta = Dr.deltas[((size_t)iat_next * Dr.nw + iw) * 3 + d] * Dr.sqrttau;
    const T rold  = J.rsoa[((size_t)iw * 3 + d) * J.npad + iat_next];
    T disp        = delta;
    if (Dr.use_drift)
    {
      // thread 0 holds the determinant gradient (every thread of the determinant group does after group_sum)
      const V gdet = d == 0 ? g[0] : (d == 1 ? g[1] : g[2]);
      if (det_grads_out && lane < 3)
        det_grads_out[3 * iw + d] = gdet;
      T gd = real_part(gdet); // convertToReal of the complex gradient (DriftModifierUNR.cpp:20-23)
      if (J.has_j2)
        gd += J.dUat[((size_t)iw * 3 + d) * J.npad + iat_next];
      if (J.has_j1)
        gd += J.Grad1[((size_t)iw * 3 + d) * J.N + iat_next];
      const T g0 = __shfl_sync(0xffffffffu, gd, 0), g1 = __shfl_sync(0xffffffffu, gd, 1),
              g2 = __shfl_sync(0xffffffffu, gd, 2);
      const T gv[3] = {g0, g1, g2};
      T dr[3];
      get_drift<T>(Dr.tauovermass, gv, dr);
      disp = (d == 0 ? dr[0] : (d == 1 ? dr[1] : dr[2])) + delta;
    }
    if (lane < 3)
    {
      Dr.drifts[3 * iw + d]    = disp;
      Dr.delta_cur[3 * iw + d] = delta;
      J.newpos[3 * iw + d]     = rold + disp;
      s_newpos[d]              = rold + disp;
    }
  }
  // Jastrow sums at the proposed position (TwoBodyJastrow::mw_ratioGrad, J1 ratioGrad), fused into the tail of the
  // boundary: a separate kernel beside the spline gather is starved by the gather's register footprint and costs more
  // than these few microseconds of the whole CTA
  if (Dr.fuse_jastrow && part2 && !twf_grads_out && (J.has_j2 || J.has_j1))
  {
    __syncthreads();
    const T pos[3] = {s_newpos[0], s_newpos[1], s_newpos[2]};
    jastrow_move_body<T, false>(J, iw, iat_next, pos, jred, jlist);
  }
}

// kinetic energy and log psi per walker: ke = -1/2 sum_i (L_i + G_i.G_i)
// complex G, L: real(CplxDot(G,G) + CplxSum(L)) = sum(re^2 - im^2) + sum(re L)  (BareKineticEnergy.cpp:114)
__device__ __forceinline__ double ke_term(const float l, const float* g)
{
  return (double)l + (double)g[0] * g[0] + (double)g[1] * g[1] + (double)g[2] * g[2];
}
__device__ __forceinline__ double ke_term(const double l, const double* g)
{
  return l + g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
}
template<typename T>
__device__ __forceinline__ double ke_term(const cx<T>& l, const cx<T>* g)
{
  double s = (double)l.re;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    s += (double)g[d].re * g[d].re - (double)g[d].im * g[d].im;
  return s;
}
template<typename V>
__global__ void __launch_bounds__(256) ke_kernel(int N, const V* Gd, const V* Ld, double* ke)
{
  __shared__ double red[32];
  const int iw = blockIdx.x, tid = threadIdx.x;
  double acc[1] = {0.0};
  for (int i = tid; i < N; i += blockDim.x)
    acc[0] += ke_term(Ld[(size_t)iw * N + i], Gd + ((size_t)iw * N + i) * 3);
  block_sum<double, 1>(acc, red);
  if (tid == 0)
    ke[iw] = -0.5 * acc[0];
}

// ------------------------------------------------------------------------------------------------------------
static int env_flag(const char* name, int dflt)
{
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

// ---- residency budget of the persistent walker-segment kernel (segment.cuh).  Its CTAs wait on one another (RNG order),
// so every CTA of a launch must be resident, and crowds of one device launch concurrently from their own streams: two
// half-resident launches would wait on each other's slots for ever.  Every crowd therefore reserves, per device, the SM
// share of ALL its walkers (walkers / CTAs-per-SM) when its driver is initialised; a crowd that does not fit any more
// runs the two-kernel path.
struct SegBudget
{
  static std::mutex& mu()
  {
    static std::mutex m;
    return m;
  }
  static std::map<int, double>& used()
  {
    static std::map<int, double> u;
    return u;
  }
  static bool reserve(int dev, double sms, int sm_count)
  {
    std::lock_guard<std::mutex> lock(mu());
    double& u = used()[dev];
    if (u + sms > (double)sm_count + 1e-9)
      return false;
    u += sms;
    return true;
  }
  static void release(int dev, double sms)
  {
    std::lock_guard<std::mutex> lock(mu());
    double& u = used()[dev];
    u         = std::max(0.0, u - sms);
  }
};

template<typename T, typename V>
struct Crowd : CrowdBase
{
  using DV                      = typename value_traits<V>::dbl_t; // FP64 counterpart of the determinant value type
  static constexpr bool cplx    = value_traits<V>::is_complex;
  static constexpr int ncomp    = value_traits<V>::ncomp;
  qmcb_system sys;
  int nw = 0, N = 0, k = 1;
  int cap = 0; // walkers the buffers were sized for; nw (<= cap) is the number of live walkers (DMC population)
  int nel[2], first[2], lda[2];
  int nmax = 0;
  size_t npad = 0;
  SplineSPOBase* spo[2];
  cudaStream_t st = nullptr, st2 = nullptr, st3 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_rng = nullptr, ev_rng_done = nullptr;
  cublasHandle_t blas = nullptr;

  // determinants
  DetDev<V> det[2];
  DevBuf<V> Ainv[2], GL[2], U[2], Vb[2], Binv[2], wvec[2], invRow[2], tempMat[2], Up[2];
  DevBuf<int> list[2];
  DevBuf<double> logdet[2];
  int rg_cap = 1, rg_nparts = 1; // slots per walker in `rg` / slots filled by the last evaluation
  int delay_count[2] = {0, 0};
  int invrow_id[2]   = {-1, -1};
  // jastrow / particle set
  JastrowDev<T> jas;
  DevBuf<T> rsoa, newpos, rows, cur_allu, j2_vgl, Uat, dUat, d2Uat, j1_cur, Vat, Grad1, Lap1, ion_rsoa, fcoefs;
  DevBuf<int> ion_grp;
  DevBuf<double> j2_log, j1_log;
  // shared scratch
  DevBuf<V> phi_vgl, rg, det_grads, grads_tmp, Gd, Ld;
  DevBuf<T> displ;
  DevBuf<DV> ratios_d;
  DevBuf<double> ke_d;
  DevBuf<unsigned char> accepted;
  DevBuf<unsigned> err_word; // crowd-wide error bits set by the kernels (det.cuh: QMCB_ERR_*)
  PinBuf<unsigned char> h_acc;
  PinBuf<V> h_t;  // gradients / displacements staging (displacements use the second half, as T)
  PinBuf<DV> h_d;
  // driver
  DriverDev<T> drv;
  RngDev rng;
  DevBuf<T> deltas, drifts, delta_cur, rr_acc, rr_prop;
  DevBuf<uint32_t> rng_state, rng_ring;
  DevBuf<unsigned long long> rng_cnt, n_acc, n_rej;
  DevBuf<unsigned> rng_flags;
  DevBuf<unsigned char> gj_used; // work buffers of the blocked Gauss-Jordan inverse (inverse.cuh)
  DevBuf<int> gj_R, gj_Rinv, gj_info;
  DevBuf<DV> gj_W, gj_Xr;
  DevBuf<unsigned> ticket;   // dispatch-order walker tickets of oversubscribed boundary launches (driver.cuh)
  int boundary_slots = 0;    // boundary-kernel CTAs the device holds at once
  DevBuf<unsigned char> accept_log;
  bool vmc_ready = false, use_graph = false;
  // persistent walker-segment kernel (segment.cuh)
  bool fused = false;
  double seg_reserved_sms = 0.0;
  SegRng segrng{};
  DevBuf<unsigned> seg_flags, seg_tot_tag, seg_btot;
  DevBuf<unsigned long long> seg_tot_val;
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_logs = false;
  unsigned long long sweep_backlog = 0;
  size_t dev_bytes = 0;

  template<typename X>
  void A(DevBuf<X>& b, size_t n)
  {
    b.alloc(n);
    dev_bytes += b.bytes();
  }

  static void fill_functor(FunctorDev<T>& f, std::vector<T>& pool, const double* params, int np, double rcut, double cusp)
  {
    // ref: BsplineFunctor.h:102-131 resize/reset, arithmetic in Real = T
    const T cutoff     = (T)rcut;
    const int numCoefs = np + 4;
    const int numKnots = numCoefs - 2;
    const T DeltaR     = cutoff / (T)(numKnots - 1);
    const T DeltaRInv  = (T)(1.0 / DeltaR);
    std::vector<T> c(numCoefs, T(0)), P(params, params + np);
    c[1] = P[0];
    c[2] = P[1];
    c[0] = (T)(P[1] - 2.0 * DeltaR * (T)cusp);
    for (int i = 2; i < np; ++i)
      c[i + 1] = P[i];
    f.coefs     = reinterpret_cast<const T*>(pool.size()); // offset for now, patched after upload
    f.DeltaRInv = DeltaRInv;
    f.rcut      = cutoff;
    f.max_index = numCoefs - 4;
    pool.insert(pool.end(), c.begin(), c.end());
  }

  Crowd(const qmcb_system* s, int nw_) : sys(*s), nw(nw_)
  {
    if (nw <= 0)
      throw std::runtime_error("crowd: nw must be positive");
    cap = nw;
    N        = sys.n_up + sys.n_dn;
    k        = std::max(1, sys.delay_rank);
    nel[0]   = sys.n_up;
    nel[1]   = sys.n_dn;
    first[0] = 0;
    first[1] = sys.n_up;
    nmax     = std::max(nel[0], nel[1]);
    npad     = aligned_size<T>(N);
    for (int i = 0; i < 2; ++i)
    {
      if (!sys.spo[i])
        throw std::runtime_error("crowd: missing SPOSet handle");
      spo[i] = sys.spo[i]->impl.get();
      int dev_now = 0;
      QMCB_CUDA(cudaGetDevice(&dev_now));
      if (spo[i]->device != dev_now)
        throw std::runtime_error("crowd: the SPOSet lives on another CUDA device (tables are replicated per GPU)");
      if (spo[i]->precision != sys.precision || spo[i]->kind != (cplx ? QMCB_C2C : QMCB_R2R))
        throw std::runtime_error("crowd: SPOSet precision/kind mismatch (both determinants need tables of one kind: "
                                 "SplineR2R for real, SplineC2C for complex orbitals)");
      if (spo[i]->n_orb != nel[i])
        throw std::runtime_error("crowd: the SPOSet of each determinant must hold exactly n_el orbitals");
      if (k > std::max(1, nel[i]) && nel[i] > 0)
        throw std::runtime_error("crowd: delay_rank larger than the determinant");
    }
    QMCB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    QMCB_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    QMCB_CUDA(cudaStreamCreateWithFlags(&st3, cudaStreamNonBlocking));
    QMCB_CUDA(cudaEventCreateWithFlags(&ev_rng_done, cudaEventDisableTiming));
    QMCB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    QMCB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    QMCB_CUDA(cudaEventCreateWithFlags(&ev_rng, cudaEventDisableTiming));
    QMCB_CUBLAS(cublasCreate(&blas));
    QMCB_CUBLAS(cublasSetStream(blas, st));

    for (int s2 = 0; s2 < 2; ++s2)
    {
      const int n = nel[s2];
      lda[s2]     = (int)aligned_size<V>(n);
      A(Ainv[s2], (size_t)nw * n * lda[s2]);
      A(GL[s2], (size_t)nw * n * 4 * n);
      A(U[s2], (size_t)nw * k * n);
      A(Vb[s2], (size_t)nw * k * n);
      A(Binv[s2], (size_t)nw * k * k);
      A(wvec[s2], (size_t)nw * k);
      A(list[s2], (size_t)nw * k);
      A(invRow[s2], (size_t)nw * n);
      A(tempMat[s2], (size_t)nw * n * k);
      A(Up[s2], (size_t)nw * k * n);
      A(logdet[s2], (size_t)nw * 2);
      DetDev<V>& D = det[s2];
      D.n = n, D.lda = lda[s2], D.k = k, D.nw = nw;
      D.Ainv = Ainv[s2].p, D.GL = GL[s2].p, D.U = U[s2].p, D.V = Vb[s2].p, D.Binv = Binv[s2].p, D.wvec = wvec[s2].p;
      D.list = list[s2].p, D.invRow = invRow[s2].p, D.tempMat = tempMat[s2].p, D.Up = Up[s2].p, D.logdet = logdet[s2].p;
    }
    A(err_word, 4);
    det[0].err = det[1].err = err_word.p;
    A(phi_vgl, (size_t)5 * nw * nmax);
    rg_cap = std::max(1, std::max(spo[0]->rg_parts(), spo[1]->rg_parts()));
    A(rg, (size_t)nw * rg_cap * 4);
    A(det_grads, (size_t)nw * 3);
    A(grads_tmp, (size_t)nw * 3);
    A(displ, (size_t)nw * 3);
    A(Gd, (size_t)nw * N * 3);
    A(Ld, (size_t)nw * N);
    A(ratios_d, nw);
    A(ke_d, nw);
    A(accepted, nw);
    h_acc.alloc(4 * (size_t)nw);
    h_t.alloc(std::max<size_t>((size_t)nw * 8, 64));
    h_d.alloc(std::max<size_t>((size_t)nw * 8, 64));

    // ---- particle set + Jastrows
    std::memset(&jas, 0, sizeof(jas));
    jas.N = N, jas.npad = (int)npad, jas.n_up = sys.n_up, jas.nw = nw;
    set_cell(jas.cell, sys.lattice);
    A(rsoa, (size_t)nw * 3 * npad);
    A(newpos, (size_t)nw * 3);
    jas.rsoa = rsoa.p, jas.newpos = newpos.p;
    std::vector<T> pool;
    jas.has_j2 = sys.n_j2 > 0;
    if (jas.has_j2)
    {
      A(rows, (size_t)2 * nw * 4 * npad);
      A(j2_vgl, (size_t)nw * 5);
      A(Uat, (size_t)nw * npad);
      A(dUat, (size_t)nw * 3 * npad);
      A(d2Uat, (size_t)nw * npad);
      A(j2_log, nw);
      jas.rows = rows.p, jas.cur_allu = nullptr, jas.j2_vgl = j2_vgl.p, jas.Uat = Uat.p, jas.dUat = dUat.p;
      jas.d2Uat = d2Uat.p, jas.j2_log = j2_log.p;
      // cusp -1/4 like spin, -1/2 unlike spin (Jastrow/RadialJastrowBuilder.cpp:200-208).  An input that only gives the
      // like-spin correlation (j2_ud == NULL) uses that functor, cusp included, for every pair: TwoBodyJastrow::addFunc
      // fills all pair slots when the first (u, u) functor is added (Jastrow/TwoBodyJastrow.cpp, addFunc).
      const double* ud     = sys.j2_ud ? sys.j2_ud : sys.j2_uu;
      const double cusp_ud = sys.j2_ud ? -0.5 : -0.25;
      fill_functor(jas.F2[0], pool, sys.j2_uu, sys.n_j2, sys.j2_rcut, -0.25);
      fill_functor(jas.F2[1], pool, ud, sys.n_j2, sys.j2_rcut, cusp_ud);
      fill_functor(jas.F2[2], pool, ud, sys.n_j2, sys.j2_rcut, cusp_ud);
      fill_functor(jas.F2[3], pool, sys.j2_uu, sys.n_j2, sys.j2_rcut, -0.25);
    }
    jas.has_j1 = sys.nions > 0;
    if (jas.has_j1)
    {
      if (sys.n_ion_groups > 8)
        throw std::runtime_error("crowd: at most 8 ion groups");
      jas.nions    = sys.nions;
      jas.npad_ion = (int)aligned_size<T>(sys.nions);
      std::vector<T> ir(3 * (size_t)jas.npad_ion, T(0));
      for (int i = 0; i < sys.nions; ++i)
        for (int d = 0; d < 3; ++d)
          ir[(size_t)d * jas.npad_ion + i] = (T)sys.ion_pos[3 * i + d];
      A(ion_rsoa, ir.size());
      QMCB_CUDA(cudaMemcpy(ion_rsoa.p, ir.data(), ir.size() * sizeof(T), cudaMemcpyHostToDevice));
      A(ion_grp, sys.nions);
      QMCB_CUDA(cudaMemcpy(ion_grp.p, sys.ion_grp, sys.nions * sizeof(int), cudaMemcpyHostToDevice));
      A(j1_cur, (size_t)nw * 5);
      A(Vat, (size_t)nw * N);
      A(Grad1, (size_t)nw * 3 * N);
      A(Lap1, (size_t)nw * N);
      A(j1_log, nw);
      jas.ion_rsoa = ion_rsoa.p, jas.ion_grp = ion_grp.p, jas.j1_cur = j1_cur.p, jas.Vat = Vat.p, jas.Grad1 = Grad1.p;
      jas.Lap1 = Lap1.p, jas.j1_log = j1_log.p;
      for (int g = 0; g < sys.n_ion_groups; ++g)
        fill_functor(jas.F1[g], pool, sys.j1_params + (size_t)g * sys.n_j1, sys.n_j1, sys.j1_rcut[g], 0.0);
    }
    if (!pool.empty())
    {
      A(fcoefs, pool.size());
      QMCB_CUDA(cudaMemcpy(fcoefs.p, pool.data(), pool.size() * sizeof(T), cudaMemcpyHostToDevice));
      auto patch = [&](FunctorDev<T>& f) { f.coefs = fcoefs.p + reinterpret_cast<size_t>(f.coefs); };
      if (jas.has_j2)
        for (int i = 0; i < 4; ++i)
          patch(jas.F2[i]);
      if (jas.has_j1)
        for (int g = 0; g < sys.n_ion_groups; ++g)
          patch(jas.F1[g]);
    }
    std::memset(&drv_host, 0, sizeof(drv_host));
    drv_host.nw = nw, drv_host.N = N, drv_host.use_drift = 1, drv_host.accepted = accepted.p;
    drv_host.err = err_word.p;
    drv_host.pdl_early = (g_pdl_mode & 4) ? 1 : 0;
    drv_host.l1_prefetch = env_flag("QMCB_L1PF", 1);
    std::memset(&rng, 0, sizeof(rng));
    QMCB_CUDA(cudaDeviceSynchronize());
  }

  ~Crowd() override
  {
    try
    {
      hd_abort();
    }
    catch (...)
    {
    }
    release_segment_slots();
    if (graph_exec)
      cudaGraphExecDestroy(graph_exec);
    if (blas)
      cublasDestroy(blas);
    if (ev_fork)
      cudaEventDestroy(ev_fork);
    if (ev_join)
      cudaEventDestroy(ev_join);
    if (ev_rng)
      cudaEventDestroy(ev_rng);
    if (st)
      cudaStreamDestroy(st);
    if (st2)
      cudaStreamDestroy(st2);
    if (st3)
      cudaStreamDestroy(st3);
    if (ev_rng_done)
      cudaEventDestroy(ev_rng_done);
  }

  // LatticeAnalyzer.h:213-275: shortest equivalent basis.  One vector at a time is replaced by the shortest of the four
  // combinations a+b-c, a+c-b, b+c-a, a+b+c (and, with one vector zeroed, the pairwise sums/differences) until nothing
  // shrinks.  Done in T like DTD_BConds<T, 3, PPPG + SOA_OFFSET> (ParticleBConds3DSoa.h:339-347).
  static bool found_shorter_base(T rb[3][3])
  {
    const T eps = T(10) * std::numeric_limits<T>::epsilon();
    auto dot3   = [](const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    int imax    = 0;
    T r2max     = dot3(rb[0], rb[0]);
    for (int i = 1; i < 3; ++i)
    {
      const T r2 = dot3(rb[i], rb[i]);
      if ((r2 - r2max) > eps)
      {
        r2max = r2;
        imax  = i;
      }
    }
    const T rmax = std::sqrt(r2max);
    const T tol  = T(4) * rmax * eps;
    T cand[4][3];
    for (int d = 0; d < 3; ++d)
    {
      cand[0][d] = rb[0][d] + rb[1][d] - rb[2][d];
      cand[1][d] = rb[0][d] + rb[2][d] - rb[1][d];
      cand[2][d] = rb[1][d] + rb[2][d] - rb[0][d];
      cand[3][d] = rb[0][d] + rb[1][d] + rb[2][d];
    }
    for (int i = 0; i < 4; ++i)
      if ((dot3(cand[i], cand[i]) - r2max) < -tol)
      {
        for (int d = 0; d < 3; ++d)
          rb[imax][d] = cand[i][d];
        return true;
      }
    return false;
  }
  static void find_reduced_basis(T rb[3][3])
  {
    for (int count = 0; count < 10000; ++count)
    {
      T saved[3][3];
      std::memcpy(saved, rb, sizeof(saved));
      bool changed = false;
      for (int i = 0; i < 3; ++i)
      {
        rb[i][0] = rb[i][1] = rb[i][2] = T(0);
        changed                        = found_shorter_base(rb);
        for (int d = 0; d < 3; ++d)
          rb[i][d] = saved[i][d];
        if (changed)
          break;
      }
      if (!changed && !found_shorter_base(rb))
        return;
    }
    throw std::runtime_error("crowd: reduced basis not found in the allowed number of iterations; check the unit cell");
  }

  // minimum-image data of the cell.  Orthorhombic cells: box lengths (DTD_BConds PPPO, ParticleBConds3DSoa.h:111-139).
  // General cells: rows, inverse and the 8 corner shifts are all taken from the REDUCED basis (:339-386) -- with the raw
  // rows of a skewed cell the floor + 8-corner search can miss the nearest image.
  static void set_cell(CellDev<T>& C, const double R[9])
  {
    C.ortho = (R[1] == 0 && R[2] == 0 && R[3] == 0 && R[5] == 0 && R[6] == 0 && R[7] == 0) ? 1 : 0;
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
        R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det == 0)
      throw std::runtime_error("crowd: singular lattice");
    T rb[3][3];
    for (int i = 0; i < 3; ++i)
      for (int d = 0; d < 3; ++d)
        rb[i][d] = (T)R[3 * i + d];
    if (!C.ortho)
      find_reduced_basis(rb);
    // inverse(Tensor<T, 3>) in T (OhmmsPETE/TensorOps.h:906-923)
    const T* a   = &rb[0][0];
    const T detT = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    const T vinv = T(1) / detT;
    C.g[0] = vinv * (a[4] * a[8] - a[5] * a[7]);
    C.g[1] = vinv * (a[7] * a[2] - a[8] * a[1]);
    C.g[2] = vinv * (a[1] * a[5] - a[2] * a[4]);
    C.g[3] = vinv * (a[5] * a[6] - a[3] * a[8]);
    C.g[4] = vinv * (a[8] * a[0] - a[6] * a[2]);
    C.g[5] = vinv * (a[2] * a[3] - a[0] * a[5]);
    C.g[6] = vinv * (a[3] * a[7] - a[4] * a[6]);
    C.g[7] = vinv * (a[6] * a[1] - a[7] * a[0]);
    C.g[8] = vinv * (a[0] * a[4] - a[1] * a[3]);
    for (int i = 0; i < 9; ++i)
      C.r[i] = a[i];
    for (int d = 0; d < 3; ++d)
    {
      const double len = std::sqrt(R[3 * d] * R[3 * d] + R[3 * d + 1] * R[3 * d + 1] + R[3 * d + 2] * R[3 * d + 2]);
      C.L[d]           = (T)len;
      C.Linv[d]        = (T)(1.0 / len);
      const T a0 = rb[0][d], a1 = rb[1][d], a2 = rb[2][d];
      C.corners[d][0] = T(0);
      C.corners[d][1] = T(-1) * a0;
      C.corners[d][2] = T(-1) * a1;
      C.corners[d][3] = T(-1) * a2;
      C.corners[d][4] = T(-1) * (a0 + a1);
      C.corners[d][5] = T(-1) * (a0 + a2);
      C.corners[d][6] = T(-1) * (a1 + a2);
      C.corners[d][7] = T(-1) * (a0 + a1 + a2);
    }
  }

  cudaStream_t stream() override { return st; }
  void sync() override
  {
    hd_abort(); // (a resident kernel waiting for the host would never let the stream drain)
    QMCB_CUDA(cudaStreamSynchronize(st));
  }
  // host-driven move loop: two device round trips per move, so the wake-up latency of a blocking synchronize matters;
  // poll the stream instead (one host thread per crowd, VMCBatched.cpp:348)
  // Polling with cudaStreamQuery takes the driver's context lock on every call: with one polling thread per crowd the
  // launches of the other crowds queue behind the pollers (measured: 2 crowds overlap, 4 and 8 do not).  Default: a
  // stream memory operation (cuStreamWriteValue32) publishes a sequence number in pinned host memory when the stream
  // reaches it and the host thread spins on that word -- one driver call per round trip, none while waiting.
  // QMCB_SYNC=query keeps the stream-query poll.
  typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static StreamWriteValue32Fn stream_write_fn()
  {
    static StreamWriteValue32Fn fn = [] {
      const char* e = std::getenv("QMCB_SYNC");
      if (e && std::string(e) == "query")
        return (StreamWriteValue32Fn) nullptr;
      void* p = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &qres) != cudaSuccess || !p ||
          qres != cudaDriverEntryPointSuccess)
        return (StreamWriteValue32Fn) nullptr;
      return reinterpret_cast<StreamWriteValue32Fn>(p);
    }();
    return fn;
  }
  PinBuf<uint32_t> sync_word;
  uint32_t sync_seq = 0;
  void spin_sync()
  {
    StreamWriteValue32Fn wr = stream_write_fn();
    if (wr)
    {
      if (!sync_word.p)
      {
        sync_word.alloc(16);
        sync_word.p[0] = 0;
      }
      ++sync_seq;
      if (wr((CUstream)st, (CUdeviceptr)(uintptr_t)sync_word.p, sync_seq, 0) == CUDA_SUCCESS)
      {
        volatile uint32_t* w = sync_word.p;
        unsigned long long spins = 0;
        while (*w != sync_seq)
        {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
          if ((++spins & 0xfffffull) == 0) // every ~million spins: surface a sticky device error instead of hanging
          {
            const cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady)
              QMCB_CUDA(e);
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return;
      }
    }
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady)
    {
    }
    QMCB_CUDA(e);
  }
  size_t device_bytes() const override { return dev_bytes; }
  bool is_complex() const override { return cplx; }
  int spin_of(int iat) const { return iat < sys.n_up ? 0 : 1; }
  static int blocks(int n, int tpb) { return (n + tpb - 1) / tpb; }
  void check_iat(int iat) const
  {
    if (iat < 0 || iat >= N)
      throw std::runtime_error("electron index out of range");
  }
  void check_row(int spin, int row) const
  {
    if (spin < 0 || spin > 1 || row < 0 || row >= nel[spin])
      throw std::runtime_error("determinant row out of range");
  }

  // ---------------------------------------------------------------- numerical guards
  // The reference throws from the host the moment a ratio is NaN (NaNguard::checkOneParticleRatio, TrialWaveFunction.cpp:
  // 473,508,549) or a move is accepted with a zero determinant ratio (DiracDeterminantBatched.cpp:494-500).  Kernels that
  // decide on the device leave a bit in err_word instead; every synchronising entry point turns it into the same exception.
  void check_device_errors(const char* where)
  {
    unsigned bits = 0;
    QMCB_CUDA(cudaMemcpyAsync(&bits, err_word.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaStreamSynchronize(st));
    if (!bits)
      return;
    QMCB_CUDA(cudaMemsetAsync(err_word.p, 0, sizeof(unsigned), st));
    std::string msg;
    if (bits & QMCB_ERR_NAN_RATIO)
      msg += std::string("NaNguard::checkOneParticleRatio error message: ") + where + ": a one-particle ratio is NaN. ";
    if (bits & QMCB_ERR_ZERO_RATIO_ACCEPTED)
      msg += std::string("mw_accept_rejectMove (") + where + "): det.curRatio is 0 for an accepted move! Report a bug.";
    throw std::runtime_error(msg);
  }
  std::vector<unsigned char> ratio_zero; // host-driven path: walkers whose last determinant-times-Jastrow ratio was 0
  void guard_ratios(const double* ratios, const char* where)
  {
    ratio_zero.assign(nw, 0);
    for (int iw = 0; iw < nw; ++iw)
    {
      double nrm = 0;
      for (int c2 = 0; c2 < ncomp; ++c2)
        nrm += ratios[(size_t)iw * ncomp + c2] * ratios[(size_t)iw * ncomp + c2];
      if (std::isnan(nrm))
        throw std::runtime_error(std::string("NaNguard::checkOneParticleRatio error message: ") + where + " walker " +
                                 std::to_string(iw) + ": ratio is NaN");
      ratio_zero[iw] = nrm == 0.0;
    }
  }
  void guard_accept(const uint8_t* acc)
  {
    if ((int)ratio_zero.size() != nw)
      return;
    for (int iw = 0; iw < nw; ++iw)
      if (acc[iw] && ratio_zero[iw])
      {
        ratio_zero.clear();
        throw std::runtime_error("mw_accept_rejectMove: curRatio is 0 for the accepted move of walker " + std::to_string(iw) +
                                 "! Report a bug.");
      }
    ratio_zero.clear();
  }

  // ---------------------------------------------------------------- positions
  void set_positions(const double* R) override
  {
    flush_pending();
    std::vector<T> h((size_t)nw * 3 * npad, T(0));
    for (int iw = 0; iw < nw; ++iw)
      for (int i = 0; i < N; ++i)
        for (int d = 0; d < 3; ++d)
          h[((size_t)iw * 3 + d) * npad + i] = (T)R[((size_t)iw * N + i) * 3 + d];
    QMCB_CUDA(cudaMemcpyAsync(rsoa.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    sync();
  }
  void get_positions(double* R) override
  {
    flush_pending();
    std::vector<T> h((size_t)nw * 3 * npad);
    QMCB_CUDA(cudaMemcpyAsync(h.data(), rsoa.p, h.size() * sizeof(T), cudaMemcpyDeviceToHost, st));
    sync();
    for (int iw = 0; iw < nw; ++iw)
      for (int i = 0; i < N; ++i)
        for (int d = 0; d < 3; ++d)
          R[((size_t)iw * N + i) * 3 + d] = (double)h[((size_t)iw * 3 + d) * npad + i];
  }

  // ---------------------------------------------------------------- determinant engine
  void launch_prepare(int spin, int row, V* grads)
  {
    const DetDev<V>& D = det[spin];
    const size_t smem  = (size_t)(D.n + 2 * D.k) * sizeof(V);
    det_prepare_row_kernel<V><<<nw, DET_TPB, smem, st>>>(D, row, delay_count[spin], grads);
    QMCB_LAUNCH_CHECK();
    invrow_id[spin] = row;
  }
  void ensure_row(int spin, int row)
  {
    if (invrow_id[spin] != row)
      launch_prepare(spin, row, nullptr);
  }
  template<int KD, int KC, int STAGES>
  void launch_flush_dmma_as(const DetDev<V>& D, int c, bool up_ready)
  {
    constexpr size_t smem = wb64::smem_bytes<V, KD, KC, STAGES>();
    ensure_dynamic_smem(wb64::woodbury_flush_dmma_kernel<V, KD, KC, STAGES>, smem);
    const int n = D.n;
    // U'[c x n] = Binv[c x c] * V[c x n]  (0.8 MFLOP per walker at a64; the fused kernel reads it back from L2)
    static const int up_simt = env_flag("QMCB_UP_SIMT", 0);
    if (!up_ready)
    {
      if (up_simt)
        wb64::binv_v_kernel<V><<<dim3(blocks(n, 128), nw, blocks(c, 8)), 128, 0, st>>>(D, c);
      else
      {
        constexpr size_t smem_b = (size_t)KD * (KD + 4) * sizeof(V);
        ensure_dynamic_smem(wb64::binv_v_dmma_kernel<V, KD>, smem_b);
        wb64::binv_v_dmma_kernel<V, KD><<<nw, wb64::TPB, smem_b, st>>>(D, c);
      }
      QMCB_LAUNCH_CHECK();
    }
    static const int split_env = [] {
      const char* e = std::getenv("QMCB_DMMA_SPLIT");
      return e ? std::atoi(e) : 0;
    }();
    const int ntiles = (n + wb64::RT - 1) / wb64::RT;
    const int split  = split_env > 0 ? std::min(split_env, ntiles) : ntiles;
    wb64::woodbury_flush_dmma_kernel<V, KD, KC, STAGES><<<dim3(split, nw), wb64::TPB, smem, st>>>(D, c);
    QMCB_LAUNCH_CHECK();
  }
  bool launch_flush_dmma(const DetDev<V>& D, int c, bool up_ready)
  {
    if constexpr (std::is_same<V, double>::value)
    {
      if (wb64::eligible<V>(D.n, D.k, c, 32))
        launch_flush_dmma_as<32, 32, 4>(D, c, up_ready);
      else if (wb64::eligible<V>(D.n, D.k, c, 64))
        launch_flush_dmma_as<64, 32, 4>(D, c, up_ready);
      else
        return false;
      return true;
    }
    else if constexpr (std::is_same<V, cx<double>>::value)
    {
      if (wb64::eligible<V>(D.n, D.k, c, 32))
        launch_flush_dmma_as<32, 16, 4>(D, c, up_ready);
      else if (wb64::eligible<V>(D.n, D.k, c, 64))
        launch_flush_dmma_as<64, 16, 4>(D, c, up_ready);
      else
        return false;
      return true;
    }
    else
      return false;
  }
  // up_ready: U' = Binv V is already in D.Up (written by the walker-segment kernel)
  void launch_flush(int spin, bool up_ready = false)
  {
    const int c = delay_count[spin];
    if (c == 0)
      return;
    const DetDev<V>& D = det[spin];
    const int n        = D.n;
    if constexpr (std::is_same<V, float>::value)
    {
      // tcgen05 / TMEM flush (woodbury_tc5.cuh); QMCB_FLUSH=mma or simt selects the older paths
      static const int flush_mode = [] {
        const char* e = std::getenv("QMCB_FLUSH");
        return !e ? 0 : (std::string(e) == "tc5" ? 0 : (std::string(e) == "mma" ? 1 : 2));
      }();
      if (flush_mode == 0 && wb5::eligible(n, D.k, c))
      {
        const size_t smem5 = wb5::smem_bytes(n);
        ensure_dynamic_smem(wb5::woodbury_flush_tc5_kernel, 200 * 1024);
        wb5::woodbury_flush_tc5_kernel<<<dim3((n + wb5::TM - 1) / wb5::TM, nw), wb5::TPB, smem5, st>>>(D, c, up_ready ? 1 : 0);
        QMCB_LAUNCH_CHECK();
        delay_count[spin] = 0;
        invrow_id[spin]   = -1;
        return;
      }
      if (flush_mode <= 1)
      {
      // one-pass tensor-core flush (woodbury.cuh) whenever U, U' and a 64-row tile of Ainv fit in shared memory
      const size_t smem = wb::smem_bytes_f32(n);
      if (c <= wb::KD && n % 4 == 0 && smem <= 227 * 1024)
      {
        ensure_dynamic_smem(wb::woodbury_flush_tf32_kernel, 227 * 1024);
        const int ntiles = (n + wb::RT - 1) / wb::RT;
        const int split  = std::min(ntiles, 2);
        wb::woodbury_flush_tf32_kernel<<<dim3(nw, split), wb::TPB, smem, st>>>(D, c);
        QMCB_LAUNCH_CHECK();
        delay_count[spin] = 0;
        invrow_id[spin]   = -1;
        return;
      }
      }
    }
    if constexpr (std::is_same<V, double>::value || std::is_same<V, cx<double>>::value)
    {
      // full precision: one pass over Ainv on the FP64 tensor pipe (woodbury_dmma.cuh); QMCB_FLUSH=simt keeps the GEMMs
      static const bool use_dmma = [] {
        const char* e = std::getenv("QMCB_FLUSH");
        return !e || std::string(e) != "simt";
      }();
      if (use_dmma && launch_flush_dmma(D, c, up_ready))
      {
        delay_count[spin] = 0;
        invrow_id[spin]   = -1;
        return;
      }
    }
    // tempMat[n x c] = Ainv[n x n] * U^T, with the -1 fix-up (applyW) fused
    gemm_batched_kernel<V, true, true><<<dim3(blocks(c, 64), blocks(n, 64), nw), 256, 0, st>>>(
        n, c, n, V(1), D.Ainv, D.lda, (size_t)n * D.lda, D.U, n, (size_t)D.k * n, V(0), D.tempMat, D.k, (size_t)n * D.k,
        D.list, D.k);
    QMCB_LAUNCH_CHECK();
    // Up[c x n] = Binv[c x c] * V[c x n]
    gemm_batched_kernel<V, false, false><<<dim3(blocks(n, 64), blocks(c, 64), nw), 256, 0, st>>>(
        c, n, c, V(1), D.Binv, D.k, (size_t)D.k * D.k, D.V, n, (size_t)D.k * n, V(0), D.Up, n, (size_t)D.k * n, nullptr, 0);
    QMCB_LAUNCH_CHECK();
    // Ainv -= tempMat * Up
    gemm_batched_kernel<V, false, false><<<dim3(blocks(n, 64), blocks(n, 64), nw), 256, 0, st>>>(
        n, n, c, V(-1), D.tempMat, D.k, (size_t)n * D.k, D.Up, n, (size_t)D.k * n, V(1), D.Ainv, D.lda, (size_t)n * D.lda,
        nullptr, 0);
    QMCB_LAUNCH_CHECK();
    delay_count[spin] = 0;
    invrow_id[spin]   = -1;
  }
  void launch_accept(int spin, int row, const unsigned char* acc_dev, const V* rg_dev, const V* phi_dev)
  {
    const DetDev<V>& D = det[spin];
    const size_t smem  = (size_t)(D.n + 2 * D.k) * sizeof(V);
    det_accept_kernel<V><<<nw, DET_TPB, smem, st>>>(D, row, delay_count[spin], acc_dev, rg_dev, rg_nparts, phi_dev);
    QMCB_LAUNCH_CHECK();
    delay_count[spin]++;
    invrow_id[spin] = -1;
    if (delay_count[spin] == k)
      launch_flush(spin);
  }
  void launch_spline(int spin, int mode, const void* invrow, size_t ld, void* phi, void* rgp, cudaStream_t s)
  {
    spo[spin]->evaluate_dev(mode, nw, newpos.p, invrow, ld, nullptr, phi, rgp, s);
    if (rgp)
      rg_nparts = spo[spin]->rg_parts();
  }

  void det_eval_grad(int spin, int row, void* grads) override
  {
    flush_pending();
    check_row(spin, row);
    launch_prepare(spin, row, det_grads.p);
    QMCB_CUDA(cudaMemcpyAsync(h_t.p, det_grads.p, (size_t)nw * 3 * sizeof(V), cudaMemcpyDeviceToHost, st));
    sync();
    std::memcpy(grads, h_t.p, (size_t)nw * 3 * sizeof(V));
  }
  void det_get_inv_row(int spin, int row, const void** dev, size_t* ld, void* host) override
  {
    flush_pending();
    check_row(spin, row);
    ensure_row(spin, row);
    if (dev)
      *dev = invRow[spin].p;
    if (ld)
      *ld = det[spin].n;
    if (host)
    {
      QMCB_CUDA(cudaMemcpyAsync(host, invRow[spin].p, (size_t)nw * det[spin].n * sizeof(V), cudaMemcpyDeviceToHost, st));
      sync();
    }
  }
  void det_ratio_grad(int spin, int row, void* ratios, void* grads, bool from_phi) override
  {
    QMCB_NVTX("DiracDeterminantBatched::ratio");
    flush_pending();
    check_row(spin, row);
    ensure_row(spin, row);
    if (from_phi)
    {
      det_ratio_from_phi_kernel<V><<<nw, DET_TPB, 0, st>>>(det[spin], phi_vgl.p, rg.p);
      QMCB_LAUNCH_CHECK();
      rg_nparts = 1;
    }
    else
      launch_spline(spin, MODE_VGL, invRow[spin].p, det[spin].n, phi_vgl.p, rg.p, st);
    std::vector<V> parts((size_t)nw * rg_nparts * 4);
    QMCB_CUDA(cudaMemcpyAsync(parts.data(), rg.p, parts.size() * sizeof(V), cudaMemcpyDeviceToHost, st));
    sync();
    V* r = static_cast<V*>(ratios);
    V* g = static_cast<V*>(grads);
    for (int iw = 0; iw < nw; ++iw)
    {
      V q[4];
      sum_rg_parts<V, 4>(parts.data(), iw, rg_nparts, q);
      r[iw] = q[0];
      if (g)
        for (int d = 0; d < 3; ++d)
          g[3 * iw + d] = q[1 + d] / q[0]; // SPOSet.cpp:171 grads = dot / ratio
    }
  }
  // pinned staging rotates over 4 slots: every move contains at least one stream sync after its accept call
  // (calc_ratio_grad / eval_grad of the next electron), so a slot is long drained when it comes round again
  int acc_slot = 0;
  // host-driven trial-wavefunction path: the boundary kernel reads the flags straight from the pinned staging slot
  // (a few hundred bytes over PCIe inside the kernel instead of a separate copy operation in front of it)
  const unsigned char* pending_flags = nullptr;
  void stage_flags(const uint8_t* acc)
  {
    unsigned char* h = h_acc.p + (size_t)(acc_slot++ & 3) * cap;
    std::memcpy(h, acc, nw);
    pending_flags = h;
  }
  void upload_flags(const uint8_t* acc)
  {
    unsigned char* h = h_acc.p + (size_t)(acc_slot++ & 3) * cap;
    std::memcpy(h, acc, nw);
    QMCB_CUDA(cudaMemcpyAsync(accepted.p, h, nw, cudaMemcpyHostToDevice, st));
  }
  void det_accept_reject(int spin, int row, const uint8_t* acc) override
  {
    QMCB_NVTX("DiracDeterminantBatched::update");
    flush_pending();
    check_row(spin, row);
    upload_flags(acc);
    launch_accept(spin, row, accepted.p, rg.p, phi_vgl.p);
    sync(); // h_acc is reused by the next call
    check_device_errors("DiracDeterminantBatched::mw_accept_rejectMove");
  }
  void det_complete_updates(int spin, void* psiMinv, double* logdet_h) override
  {
    QMCB_NVTX("DiracDeterminantBatched::update");
    flush_pending();
    if (spin < 0 || spin > 1)
      throw std::runtime_error("bad spin");
    launch_flush(spin);
    const int n = nel[spin];
    if (psiMinv)
      QMCB_CUDA(cudaMemcpy2DAsync(psiMinv, (size_t)n * sizeof(V), Ainv[spin].p, (size_t)lda[spin] * sizeof(V),
                                  (size_t)n * sizeof(V), (size_t)nw * n, cudaMemcpyDeviceToHost, st));
    if (logdet_h)
      QMCB_CUDA(cudaMemcpyAsync(logdet_h, logdet[spin].p, (size_t)nw * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync();
  }
  int det_delay_count(int spin) override
  {
    if (hd.active && spin == hd.spin)
      return hd.c0 + (hd.iat - hd.iat0); // (moves the resident kernel has been told to apply)
    return delay_count[spin];
  }
  // measurement hook: `reps` back-to-back mw_updateInvMat launches with `c` pending slots, CUDA events on the crowd stream.
  // The delay buffers hold whatever the last moves left there, so the flushes are arithmetic on stale data: the inverse
  // is saved before and restored afterwards and the crowd is left exactly as a qmcb_twf_mw_complete_updates call leaves it.
  void det_time_update_inv_mat(int spin, int c, int reps, double* us_per_call) override
  {
    flush_pending();
    if (spin < 0 || spin > 1 || c < 1 || c > k || reps < 1)
      throw std::runtime_error("det_time_update_inv_mat: need spin in {0, 1}, 1 <= delay_count <= delay_rank and reps >= 1");
    twf_complete_updates();
    DevBuf<V> saved;
    saved.alloc(Ainv[spin].n, false);
    QMCB_CUDA(cudaMemcpyAsync(saved.p, Ainv[spin].p, Ainv[spin].bytes(), cudaMemcpyDeviceToDevice, st));
    cudaEvent_t e0, e1;
    QMCB_CUDA(cudaEventCreate(&e0));
    QMCB_CUDA(cudaEventCreate(&e1));
    delay_count[spin] = c;
    launch_flush(spin); // warm-up (first-use attribute calls, instruction cache)
    QMCB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i)
    {
      delay_count[spin] = c;
      launch_flush(spin);
    }
    QMCB_CUDA(cudaEventRecord(e1, st));
    QMCB_CUDA(cudaMemcpyAsync(Ainv[spin].p, saved.p, Ainv[spin].bytes(), cudaMemcpyDeviceToDevice, st));
    QMCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    QMCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *us_per_call = 1e3 * ms / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    delay_count[spin] = 0;
    invrow_id[spin]   = -1;
    sync();
  }
  void det_set_phi_vgl(int spin, const void* phi) override
  {
    const int n = nel[spin];
    // host layout [5][nw][n] -> device layout [5][nw][n] with the same n (phi_vgl is sized for nmax but indexed with D.n)
    QMCB_CUDA(cudaMemcpyAsync(phi_vgl.p, phi, (size_t)5 * nw * n * sizeof(V), cudaMemcpyHostToDevice, st));
    sync();
  }

  // FP64 inversion of the transposed matrices in AT (column-major psiM), result into Ainv, log-determinants
  // (always FullPrecValueType like DiracMatrixInverterCUDA::mw_invertTranspose, DiracMatrixInverterCUDA.hpp:306-369).
  // Default: the blocked Gauss-Jordan inverse of csrc/inverse.cuh (panel kernel + DMMA rank-b updates over the whole
  // batch).  QMCB_LU=cublas selects cublas<t>getrfBatched + getriBatched, the reference's routines
  // (detail/CUDA/cuBLAS_LU.cu:61-210) -- kept as the baseline bench.py times beside it, and for matrices whose column
  // panel does not fit in shared memory at the narrowest panel width.
  static int gj_panel_width(int n)
  {
    int dev = 0, cap = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    for (int b : {32, 16, 8})
      if (gj::panel_smem<DV>(n, b) <= (size_t)cap)
        return b;
    return 0;
  }
  void invert_from_AT(int spin, DevBuf<DV>& AT, int method = 0 /* 0: default, 1: cuBLAS, 2: Gauss-Jordan */)
  {
    QMCB_NVTX("DiracDeterminantBatched::inverse");
    static const bool force_cublas = [] {
      const char* e = std::getenv("QMCB_LU");
      return e && std::string(e) == "cublas";
    }();
    const int n  = nel[spin];
    const int BW = (method == 1 || (method == 0 && force_cublas)) ? 0 : gj_panel_width(n);
    if (BW == 0)
    {
      invert_from_AT_cublas(spin, AT);
      return;
    }
    // work buffers of the inversion: kept by the crowd (an allocation inside every recompute costs milliseconds)
    DevBuf<unsigned char>& used = gj_used;
    DevBuf<int>&R = gj_R, &Rinv = gj_Rinv, &info = gj_info;
    DevBuf<DV>&W = gj_W, &Xr = gj_Xr;
    auto fit = [](auto& b, size_t need) {
      if (b.n < need)
        b.alloc(need, false);
    };
    fit(used, (size_t)cap * n);
    fit(R, (size_t)cap * n);
    fit(Rinv, (size_t)cap * n);
    fit(info, (size_t)cap);
    fit(W, (size_t)cap * n * BW);
    fit(Xr, (size_t)cap * BW * n);
    const size_t psm = gj::panel_smem<DV>(n, BW);
    const int nt = (n + gj::TILE - 1) / gj::TILE;
    auto update = [&](auto kern, int j0, int b) {
      const size_t usm = (size_t)(gj::TILE * (BW + 4) + 2 * BW * (gj::TILE + 4)) * sizeof(DV);
      ensure_dynamic_smem(kern, usm);
      kern<<<dim3(nt, nw), gj::TPB, usm, st>>>(AT.p, n, j0, b, W.p, Xr.p);
      QMCB_LAUNCH_CHECK();
    };
    for (int j0 = 0; j0 < n; j0 += BW)
    {
      const int b = std::min(BW, n - j0);
      auto panel = [&](auto kern) {
        ensure_dynamic_smem(kern, psm);
        kern<<<nw, gj::TPB, psm, st>>>(AT.p, n, j0, b, used.p, R.p, W.p, Xr.p, logdet[spin].p, info.p);
        QMCB_LAUNCH_CHECK();
      };
      if (BW == 32)
        panel(gj::gj_panel_kernel<DV, 32>);
      else if (BW == 16)
        panel(gj::gj_panel_kernel<DV, 16>);
      else
        panel(gj::gj_panel_kernel<DV, 8>);
      if (n > b)
      {
        if (BW == 32)
          update(gj::gj_update_kernel<DV, 32>, j0, b);
        else if (BW == 16)
          update(gj::gj_update_kernel<DV, 16>, j0, b);
        else
          update(gj::gj_update_kernel<DV, 8>, j0, b);
      }
    }
    gj::gj_perm_kernel<<<nw, 128, (size_t)(n + 15) / 16 * 16, st>>>(R.p, Rinv.p, n, logdet[spin].p);
    QMCB_LAUNCH_CHECK();
    gj::gj_store_inverse_kernel<V><<<dim3(blocks(n, 128), n, nw), 128, 0, st>>>(det[spin], AT.p, R.p, Rinv.p);
    QMCB_LAUNCH_CHECK();
    finish_inversion(spin, info);
  }
  void finish_inversion(int spin, DevBuf<int>& info)
  {
    std::vector<int> hinfo(nw);
    QMCB_CUDA(cudaMemcpyAsync(hinfo.data(), info.p, nw * sizeof(int), cudaMemcpyDeviceToHost, st));
    sync();
    for (int iw = 0; iw < nw; ++iw)
      if (hinfo[iw] != 0)
        throw std::runtime_error("matrix inversion failed (singular Slater matrix) for walker " + std::to_string(iw));
    delay_count[spin] = 0;
    invrow_id[spin]   = -1;
  }
  void invert_from_AT_cublas(int spin, DevBuf<DV>& AT)
  {
    const int n = nel[spin];
    DevBuf<DV> inv;
    DevBuf<int> piv, info;
    DevBuf<DV*> ptrs;
    inv.alloc((size_t)nw * n * n, false);
    piv.alloc((size_t)nw * n);
    info.alloc(nw);
    ptrs.alloc(2 * (size_t)nw);
    std::vector<DV*> hp(2 * (size_t)nw);
    for (int iw = 0; iw < nw; ++iw)
    {
      hp[iw]      = AT.p + (size_t)iw * n * n;
      hp[nw + iw] = inv.p + (size_t)iw * n * n;
    }
    QMCB_CUDA(cudaMemcpyAsync(ptrs.p, hp.data(), hp.size() * sizeof(DV*), cudaMemcpyHostToDevice, st));
    if constexpr (cplx)
      QMCB_CUBLAS(cublasZgetrfBatched(blas, n, reinterpret_cast<cuDoubleComplex**>(ptrs.p), n, piv.p, info.p, nw));
    else
      QMCB_CUBLAS(cublasDgetrfBatched(blas, n, ptrs.p, n, piv.p, info.p, nw));
    g_launch_count.fetch_add(1);
    det_logdet_kernel<DV><<<nw, 128, 0, st>>>(AT.p, piv.p, n, logdet[spin].p);
    QMCB_LAUNCH_CHECK();
    if constexpr (cplx)
      QMCB_CUBLAS(cublasZgetriBatched(blas, n, reinterpret_cast<cuDoubleComplex**>(ptrs.p), n, piv.p,
                                      reinterpret_cast<cuDoubleComplex**>(ptrs.p + nw), n, info.p, nw));
    else
      QMCB_CUBLAS(cublasDgetriBatched(blas, n, ptrs.p, n, piv.p, ptrs.p + nw, n, info.p, nw));
    g_launch_count.fetch_add(1);
    det_store_inverse_kernel<V><<<dim3(blocks(n, 128), n, nw), 128, 0, st>>>(det[spin], inv.p);
    QMCB_LAUNCH_CHECK();
    finish_inversion(spin, info);
  }

  void det_recompute_from_matrices(int spin, const void* psiM, const void* dpsiM, const void* d2psiM) override
  {
    const int n = nel[spin];
    const V* pm = static_cast<const V*>(psiM);
    const V* dp = static_cast<const V*>(dpsiM);
    const V* d2 = static_cast<const V*>(d2psiM);
    std::vector<DV> at((size_t)nw * n * n);
    std::vector<V> gl((size_t)nw * n * 4 * n, V(0));
    for (int iw = 0; iw < nw; ++iw)
      for (int e = 0; e < n; ++e)
        for (int j = 0; j < n; ++j)
        {
          at[((size_t)iw * n + j) * n + e] = to_dbl(pm[((size_t)iw * n + e) * n + j]);
          V* g                             = &gl[((size_t)iw * n + e) * 4 * n];
          if (dp)
            for (int d = 0; d < 3; ++d)
              g[(size_t)d * n + j] = dp[(((size_t)iw * n + e) * n + j) * 3 + d];
          if (d2)
            g[(size_t)3 * n + j] = d2[((size_t)iw * n + e) * n + j];
        }
    DevBuf<DV> AT;
    AT.alloc(at.size(), false);
    QMCB_CUDA(cudaMemcpyAsync(AT.p, at.data(), at.size() * sizeof(DV), cudaMemcpyHostToDevice, st));
    QMCB_CUDA(cudaMemcpyAsync(GL[spin].p, gl.data(), gl.size() * sizeof(V), cudaMemcpyHostToDevice, st));
    invert_from_AT(spin, AT);
  }

  // ---------------------------------------------------------------- trial wavefunction level
  void twf_recompute() override
  {
    QMCB_NVTX("TrialWaveFunction::mw_evaluateLog");
    flush_pending();
    gl_stale[0] = gl_stale[1] = false;
    for (int spin = 0; spin < 2; ++spin)
    {
      const int n = nel[spin];
      if (n == 0)
        continue;
      DevBuf<DV> AT;
      build_AT(spin, AT);
      invert_from_AT(spin, AT);
    }
    if (jas.has_j2 || jas.has_j1)
    {
      jastrow_recompute_kernel<T><<<nw, JAS_TPB, 0, st>>>(jas);
      QMCB_LAUNCH_CHECK();
    }
    sync();
  }
  // psiM of determinant `spin` at the committed positions, transposed, in double; the gradient / Laplacian rows go to GL
  void build_AT(int spin, DevBuf<DV>& AT)
  {
    const int n = nel[spin];
    AT.alloc((size_t)nw * n * n, false);
    // SPOSet::mw_evaluate_notranspose for splines = loop over electrons calling mw_evaluateVGL (BsplineSet.h:142-189)
    for (int e = 0; e < n; ++e)
    {
      const int iat = first[spin] + e;
      make_move_kernel<T><<<blocks(nw, 128), 128, 0, st>>>(jas, iat, displ_zero());
      QMCB_LAUNCH_CHECK();
      launch_spline(spin, MODE_VGL, nullptr, 0, phi_vgl.p, nullptr, st);
      det_scatter_row_kernel<V><<<dim3(blocks(n, 128), nw), 128, 0, st>>>(det[spin], e, phi_vgl.p, AT.p);
      QMCB_LAUNCH_CHECK();
    }
  }
  // measurement hook: the FP64 inversion + log-determinant of the current Slater matrices of determinant `spin`, `reps`
  // times (each on a fresh copy of psiM; the copy is outside the timed span), method 1 = cuBLAS getrf/getriBatched,
  // 2 = the blocked Gauss-Jordan kernels.  The crowd is left as qmcb_twf_mw_recompute leaves it.
  void det_time_inverse(int spin, int method, int reps, double* us_per_call) override
  {
    if (spin < 0 || spin > 1 || nel[spin] == 0 || reps < 1 || method < 1 || method > 2)
      throw std::runtime_error("det_time_inverse: need spin in {0, 1}, method in {1, 2}, reps >= 1");
    flush_pending();
    DevBuf<DV> AT, work;
    build_AT(spin, AT);
    work.alloc(AT.n, false);
    cudaEvent_t e0, e1;
    QMCB_CUDA(cudaEventCreate(&e0));
    QMCB_CUDA(cudaEventCreate(&e1));
    double total = 0.0;
    for (int i = -1; i < reps; ++i) // (i = -1: warm-up, first-use attribute calls)
    {
      QMCB_CUDA(cudaMemcpyAsync(work.p, AT.p, AT.bytes(), cudaMemcpyDeviceToDevice, st));
      QMCB_CUDA(cudaEventRecord(e0, st));
      invert_from_AT(spin, work, method);
      QMCB_CUDA(cudaEventRecord(e1, st));
      QMCB_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      QMCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (i >= 0)
        total += 1e3 * ms;
    }
    *us_per_call = total / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    twf_recompute();
  }
  const T* displ_zero()
  {
    QMCB_CUDA(cudaMemsetAsync(displ.p, 0, displ.bytes(), st));
    return displ.p;
  }

  void twf_eval_grad(int iat, double* grads) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_evalGrad");
    check_iat(iat);
    if (gl_stale[0] || gl_stale[1])
    {
      flush_pending();
      refresh_stale_gl();
    }
    if (hd_eval_grad(iat, grads))
      return;
    // [accept of the previous electron, if one is pending] + inverse row + component-summed gradient: one launch
    join_jastrow();
    // the kernel stores the gradients straight into the pinned host buffer (unified addressing: pinned allocations are
    // device-accessible): the device->host transfer needs no separate copy launch
    apply_pending(iat, h_t.p);
    spin_sync();
    widen(grads, h_t.p, 3 * (size_t)nw);
  }
  // staged values -> the caller's doubles (complex values stay interleaved re, im)
  static void widen(double* out, const V* in, size_t count)
  {
    const T* src = reinterpret_cast<const T*>(in);
    for (size_t i = 0; i < count * ncomp; ++i)
      out[i] = (double)src[i];
  }

  void ps_make_move(int iat, const double* dsp) override
  {
    QMCB_NVTX("ParticleSet::mw_makeMove");
    check_iat(iat);
    if (hd.active && hd_make_move(iat, dsp)) // (a resident kernel is started by mw_evalGrad; loops without drift launch per call)
      return;
    flush_pending();
    T* h = reinterpret_cast<T*>(h_t.p + 4 * (size_t)nw); // second half of the staging buffer
    for (int i = 0; i < 3 * nw; ++i)
      h[i] = (T)dsp[i];
    static const int host_fuse = env_flag("QMCB_HOST_FUSE", 1);
    if (host_fuse && (jas.has_j2 || jas.has_j1))
    {
      make_move_jastrow_kernel<T><<<nw, JAS_TPB, 0, st>>>(jas, iat, h); // (reads the pinned host buffer: no copy launch)
      QMCB_LAUNCH_CHECK();
      last_move_iat = iat;
      return;
    }
    make_move_kernel<T><<<blocks(nw, 128), 128, 0, st>>>(jas, iat, h);
    QMCB_LAUNCH_CHECK();
    if (jas.has_j2 || jas.has_j1)
    {
      // distance rows + Jastrow sums run on the side stream, beside the spline gather of the coming ratio call
      QMCB_CUDA(cudaEventRecord(ev_fork, st));
      QMCB_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
      jastrow_move_kernel<T, false><<<nw, JAS_TPB, 0, st2>>>(jas, iat);
      QMCB_LAUNCH_CHECK();
      QMCB_CUDA(cudaEventRecord(ev_join, st2));
      jastrow_inflight = true;
      last_move_iat    = iat;
    }
  }
  bool jastrow_inflight = false;
  int last_move_iat     = -1;
  void join_jastrow()
  {
    if (jastrow_inflight)
    {
      QMCB_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
      jastrow_inflight = false;
    }
  }

  void twf_calc_ratio_grad(int iat, double* ratios, double* grads) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_calcRatioGrad");
    check_iat(iat);
    if (hd_calc_ratio_grad(iat, ratios, grads))
    {
      guard_ratios(ratios, "TWF::mw_calcRatioGrad");
      return;
    }
    apply_pending(-1, nullptr); // (make_move already applied it; the Jastrow rows keep running beside the gather below)
    const int spin = spin_of(iat), row = iat - first[spin];
    ensure_row(spin, row);
    launch_spline(spin, MODE_VGL, invRow[spin].p, det[spin].n, phi_vgl.p, rg.p, st);
    join_jastrow();
    twf_ratio_kernel<T, V><<<blocks(nw, 128), 128, 0, st>>>(jas, iat, rg.p, rg_nparts, h_d.p, h_t.p); // (pinned host)
    QMCB_LAUNCH_CHECK();
    spin_sync();
    std::memcpy(ratios, h_d.p, (size_t)nw * sizeof(DV));
    widen(grads, h_t.p, 3 * (size_t)nw);
    guard_ratios(ratios, "TWF::mw_calcRatioGrad");
  }

  void twf_accept_reject(int iat, const uint8_t* acc, int safe_to_delay) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_accept_rejectMove");
    check_iat(iat);
    guard_accept(acc);
    if (hd_accept_reject(iat, acc, safe_to_delay))
      return;
    flush_pending();
    pending_value_only = value_only_move;
    if (value_only_move)
      gl_stale[spin_of(iat)] = true;
    value_only_move = false;
    stage_flags(acc);
    // deferred: applied together with whatever the driver asks next (normally the gradient of the next electron);
    // asynchronous like the reference's mw_accept_rejectMove ("this call may go asynchronous", TwoBodyJastrow.cpp:661)
    pending_iat = iat;
    if (!safe_to_delay)
    {
      flush_pending();
      launch_flush(spin_of(iat));
    }
  }

  void twf_complete_updates() override
  {
    QMCB_NVTX("TrialWaveFunction::mw_completeUpdates");
    flush_pending();
    launch_flush(0);
    launch_flush(1);
  }

  void twf_evaluate_gl(double* G, double* L, double* logpsi, double* ke) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_evaluateGL");
    twf_complete_updates();
    refresh_stale_gl();
    QMCB_CUDA(cudaMemsetAsync(Gd.p, 0, Gd.bytes(), st));
    QMCB_CUDA(cudaMemsetAsync(Ld.p, 0, Ld.bytes(), st));
    for (int spin = 0; spin < 2; ++spin)
      if (nel[spin] > 0)
      {
        det_compute_gl_kernel<V, V><<<dim3(nel[spin], nw), 128, 0, st>>>(det[spin], first[spin], N, Gd.p, Ld.p);
        QMCB_LAUNCH_CHECK();
      }
    if (jas.has_j2 || jas.has_j1)
    {
      jastrow_add_gl_kernel<T><<<dim3(blocks(N, 256), nw), 256, 0, st>>>(jas, reinterpret_cast<T*>(Gd.p),
                                                                         reinterpret_cast<T*>(Ld.p), ncomp);
      QMCB_LAUNCH_CHECK();
    }
    ke_kernel<V><<<nw, 256, 0, st>>>(N, Gd.p, Ld.p, ke_d.p);
    QMCB_LAUNCH_CHECK();
    std::vector<V> hg, hl;
    std::vector<double> hk(nw), l0(2 * (size_t)nw), l1(2 * (size_t)nw), lj2(nw, 0.0), lj1(nw, 0.0);
    if (G)
    {
      hg.resize((size_t)nw * N * 3);
      QMCB_CUDA(cudaMemcpyAsync(hg.data(), Gd.p, hg.size() * sizeof(V), cudaMemcpyDeviceToHost, st));
    }
    if (L)
    {
      hl.resize((size_t)nw * N);
      QMCB_CUDA(cudaMemcpyAsync(hl.data(), Ld.p, hl.size() * sizeof(V), cudaMemcpyDeviceToHost, st));
    }
    QMCB_CUDA(cudaMemcpyAsync(hk.data(), ke_d.p, nw * sizeof(double), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(l0.data(), logdet[0].p, l0.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(l1.data(), logdet[1].p, l1.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jas.has_j2)
      QMCB_CUDA(cudaMemcpyAsync(lj2.data(), j2_log.p, nw * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jas.has_j1)
      QMCB_CUDA(cudaMemcpyAsync(lj1.data(), j1_log.p, nw * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync();
    check_device_errors("TrialWaveFunction::mw_evaluateGL");
    if (G)
      widen(G, hg.data(), hg.size());
    if (L)
      widen(L, hl.data(), hl.size());
    for (int iw = 0; iw < nw; ++iw)
    {
      if (ke)
        ke[iw] = hk[iw];
      if (logpsi)
        logpsi[iw] = l0[2 * iw] + l1[2 * iw] + lj2[iw] + lj1[iw];
    }
  }

  // ---------------------------------------------------------------- ratios at virtual positions (NLPP quadrature)
  // TrialWaveFunction::mw_evaluateRatios (TrialWaveFunction.cpp:1079-1110): the product over the selected components
  // of psi(R with electron ref[i] of walker wk[i] at r_vp[i]) / psi(R).  Determinants: DiracDeterminantBatched::
  // mw_evaluateRatios (DiracDeterminantBatched.cpp:812-848) -> SPOSet::mw_evaluateDetRatios with the rows of psiMinv;
  // Jastrows: TwoBodyJastrow::mw_evaluateRatios (TwoBodyJastrow.cpp:174-210), J1OrbitalSoA::evaluateRatios.
  void twf_evaluate_ratios(int nvp, const int* wk, const int* ref, const double* r_vp, int ct, double* ratios) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_evaluateRatios");
    if (nvp < 0 || ct < 0 || ct > 2)
      throw std::runtime_error("qmcb_twf_mw_evaluate_ratios: bad arguments");
    twf_complete_updates(); // the reference evaluates the non-local pseudopotential after completeUpdates: psiMinv is current
    if (nvp == 0)
      return;
    for (int i = 0; i < nvp; ++i)
      if (wk[i] < 0 || wk[i] >= nw || ref[i] < 0 || ref[i] >= N)
        throw std::runtime_error("qmcb_twf_mw_evaluate_ratios: walker or reference particle index out of range");
    std::vector<DV> det_r(nvp, DV(1.0));
    if (ct != 2)
      for (int spin = 0; spin < 2; ++spin)
      {
        const int n = nel[spin];
        std::vector<int> sel, rowid;
        std::vector<T> pos;
        for (int i = 0; i < nvp; ++i)
          if (spin_of(ref[i]) == spin)
          {
            sel.push_back(i);
            rowid.push_back(wk[i] * n + (ref[i] - first[spin]));
            for (int d = 0; d < 3; ++d)
              pos.push_back((T)r_vp[3 * (size_t)i + d]);
          }
        if (sel.empty())
          continue;
        const int ns = (int)sel.size(), parts = spo[spin]->rg_parts();
        DevBuf<T> pos_d;
        DevBuf<int> row_d;
        DevBuf<V> rg_d;
        pos_d.alloc(pos.size(), false);
        row_d.alloc(ns, false);
        rg_d.alloc((size_t)ns * parts * 4);
        QMCB_CUDA(cudaMemcpyAsync(pos_d.p, pos.data(), pos.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        QMCB_CUDA(cudaMemcpyAsync(row_d.p, rowid.data(), ns * sizeof(int), cudaMemcpyHostToDevice, st));
        spo[spin]->evaluate_dev(MODE_V, ns, pos_d.p, Ainv[spin].p, lda[spin], row_d.p, nullptr, rg_d.p, st);
        g_launch_count.fetch_add(0);
        std::vector<V> parts_h((size_t)ns * parts * 4);
        QMCB_CUDA(cudaMemcpyAsync(parts_h.data(), rg_d.p, parts_h.size() * sizeof(V), cudaMemcpyDeviceToHost, st));
        sync();
        for (int q = 0; q < ns; ++q)
        {
          V acc4[4];
          sum_rg_parts<V, 4>(parts_h.data(), q, parts, acc4);
          det_r[sel[q]] = to_dbl(acc4[0]);
        }
      }
    std::vector<double> jas_r(nvp, 1.0);
    if (ct != 1 && (jas.has_j2 || jas.has_j1))
    {
      std::vector<T> pos(3 * (size_t)nvp);
      for (size_t i = 0; i < pos.size(); ++i)
        pos[i] = (T)r_vp[i];
      DevBuf<T> pos_d;
      DevBuf<int> wk_d, ref_d;
      DevBuf<double> out_d;
      pos_d.alloc(pos.size(), false);
      wk_d.alloc(nvp, false);
      ref_d.alloc(nvp, false);
      out_d.alloc(nvp, false);
      QMCB_CUDA(cudaMemcpyAsync(pos_d.p, pos.data(), pos.size() * sizeof(T), cudaMemcpyHostToDevice, st));
      QMCB_CUDA(cudaMemcpyAsync(wk_d.p, wk, nvp * sizeof(int), cudaMemcpyHostToDevice, st));
      QMCB_CUDA(cudaMemcpyAsync(ref_d.p, ref, nvp * sizeof(int), cudaMemcpyHostToDevice, st));
      jastrow_vp_ratio_kernel<T><<<blocks(nvp, 4), 128, 0, st>>>(jas, nvp, wk_d.p, ref_d.p, pos_d.p, out_d.p);
      QMCB_LAUNCH_CHECK();
      QMCB_CUDA(cudaMemcpyAsync(jas_r.data(), out_d.p, nvp * sizeof(double), cudaMemcpyDeviceToHost, st));
      sync();
    }
    DV* out = reinterpret_cast<DV*>(ratios);
    for (int i = 0; i < nvp; ++i)
      out[i] = det_r[i] * DV(jas_r[i]);
  }

  // ---------------------------------------------------------------- ratio-only moves (no drift)
  // TrialWaveFunction::mw_calcRatio (TrialWaveFunction.cpp:494-510) -> DiracDeterminantBatched::mw_calcRatio
  // (DiracDeterminantBatched.cpp:742-789).  The reference "temporarily" still evaluates V, G and L there; here the
  // orbital VALUES are gathered alone.  An accept that follows stores the orbital row only, so the gradient / Laplacian
  // rows of the determinant go stale (UpdateMode ORB_PBYP_RATIO) and are re-evaluated from the committed positions the
  // next time they are needed, exactly as mw_evaluateGL does in that mode (DiracDeterminantBatched.cpp:630-678).
  bool gl_stale[2] = {false, false};
  void twf_calc_ratio(int iat, double* ratios) override
  {
    QMCB_NVTX("TrialWaveFunction::mw_calcRatio");
    check_iat(iat);
    hd_abort();
    apply_pending(-1, nullptr);
    const int spin = spin_of(iat), row = iat - first[spin];
    ensure_row(spin, row);
    launch_spline(spin, MODE_V, invRow[spin].p, det[spin].n, phi_vgl.p, rg.p, st);
    join_jastrow();
    twf_ratio_kernel<T, V><<<blocks(nw, 128), 128, 0, st>>>(jas, iat, rg.p, rg_nparts, h_d.p, h_t.p); // (pinned host)
    QMCB_LAUNCH_CHECK();
    spin_sync();
    std::memcpy(ratios, h_d.p, (size_t)nw * sizeof(DV));
    guard_ratios(ratios, "TWF::mw_calcRatio");
    value_only_move = true;
  }
  bool value_only_move = false; // the orbital rows of the proposed move hold values only
  void refresh_stale_gl()
  {
    for (int spin = 0; spin < 2; ++spin)
      if (gl_stale[spin])
      {
        const int n = nel[spin];
        for (int e = 0; e < n; ++e)
        {
          make_move_kernel<T><<<blocks(nw, 128), 128, 0, st>>>(jas, first[spin] + e, displ_zero());
          QMCB_LAUNCH_CHECK();
          launch_spline(spin, MODE_VGL, nullptr, 0, phi_vgl.p, nullptr, st);
          det_scatter_gl_kernel<V><<<dim3(blocks(n, 128), nw), 128, 0, st>>>(det[spin], e, phi_vgl.p);
          QMCB_LAUNCH_CHECK();
        }
        gl_stale[spin] = false;
      }
  }

  // ---------------------------------------------------------------- component level: DT rows + J2
  void dtaa_get_temp_rows(void* out) override
  {
    flush_pending();
    if (!jas.has_j2)
      throw std::runtime_error("distance rows are only kept when a two-body Jastrow is present");
    if (last_move_iat < 0)
      throw std::runtime_error("qmcb_dtaa_get_temp_rows: no move has been proposed yet");
    // the rows are not materialised on the move path (jastrow.cuh); store them now for the last proposed move
    jastrow_move_kernel<T, true><<<nw, JAS_TPB, 0, st>>>(jas, last_move_iat);
    QMCB_LAUNCH_CHECK();
    // device [2][nw][4][npad] -> host [2][nw][4][N]
    QMCB_CUDA(cudaMemcpy2DAsync(out, (size_t)N * sizeof(T), rows.p, npad * sizeof(T), (size_t)N * sizeof(T),
                                (size_t)2 * nw * 4, cudaMemcpyDeviceToHost, st));
    sync();
  }
  void j2_ratio_grad(int iat, double* ratios, void* grads) override
  {
    flush_pending();
    check_iat(iat);
    if (!jas.has_j2)
      throw std::runtime_error("no two-body Jastrow in this crowd");
    j2_ratio_kernel<T><<<blocks(nw, 128), 128, 0, st>>>(jas, iat, ke_d.p, displ.p); // (real scratch buffers)
    QMCB_LAUNCH_CHECK();
    QMCB_CUDA(cudaMemcpyAsync(ratios, ke_d.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grads)
      QMCB_CUDA(cudaMemcpyAsync(grads, displ.p, (size_t)nw * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    sync();
  }
  void j2_accept_reject(int iat, const uint8_t* acc) override
  {
    flush_pending();
    check_iat(iat);
    upload_flags(acc);
    jastrow_accept_kernel<T><<<nw, JAS_TPB, 0, st>>>(jas, iat, accepted.p);
    QMCB_LAUNCH_CHECK();
    sync();
  }
  void j2_get_state(int iw, double* Uat_h, double* dUat_h, double* d2Uat_h) override
  {
    flush_pending();
    if (!jas.has_j2 || iw < 0 || iw >= nw)
      throw std::runtime_error("j2_get_state: bad walker or no J2");
    std::vector<T> u(npad), du(3 * npad), d2(npad);
    QMCB_CUDA(cudaMemcpyAsync(u.data(), Uat.p + (size_t)iw * npad, npad * sizeof(T), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(du.data(), dUat.p + (size_t)iw * 3 * npad, 3 * npad * sizeof(T), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(d2.data(), d2Uat.p + (size_t)iw * npad, npad * sizeof(T), cudaMemcpyDeviceToHost, st));
    sync();
    for (int i = 0; i < N; ++i)
    {
      Uat_h[i]   = u[i];
      d2Uat_h[i] = d2[i];
      for (int d = 0; d < 3; ++d)
        dUat_h[(size_t)d * N + i] = du[(size_t)d * npad + i];
    }
  }

  // ---------------------------------------------------------------- device-resident VMC driver
  void vmc_init(const qmcb_vmc_params* p) override
  {
    if (graph_exec)
    {
      cudaGraphExecDestroy(graph_exec);
      graph_exec = nullptr;
    }
    std::memset(&drv, 0, sizeof(drv));
    drv.nw = nw, drv.N = N;
    drv.err = err_word.p;
    // TauParams (QMCDrivers/TauParams.hpp:29-40), unit mass
    drv.tauovermass = (T)p->tau * (T)1.0;
    drv.oneover2tau = (T)(0.5 / drv.tauovermass);
    drv.sqrttau     = (T)std::sqrt(drv.tauovermass);
    drv.use_drift   = p->use_drift;
    drv.pdl_early   = (g_pdl_mode & 4) ? 1 : 0;
    drv.l1_prefetch = env_flag("QMCB_L1PF", 1);
    drv.dmc         = p->dmc;
    {
      const char* e    = std::getenv("QMCB_FUSE_J");
      drv.fuse_jastrow = e ? std::atoi(e) : 1;
    }
    A(rr_acc, cap);
    A(rr_prop, cap);
    drv.rr_accepted = rr_acc.p, drv.rr_proposed = rr_prop.p;
    use_graph       = p->use_cuda_graph != 0;
    A(deltas, (size_t)N * cap * 3);
    A(drifts, (size_t)cap * 3);
    A(delta_cur, (size_t)cap * 3);
    A(n_acc, cap);
    A(n_rej, cap);
    A(accept_log, (size_t)N * cap);
    drv.deltas = deltas.p, drv.drifts = drifts.p, drv.delta_cur = delta_cur.p, drv.grads_now = nullptr;
    drv.accepted = accepted.p, drv.n_accept = n_acc.p, drv.n_reject = n_rej.p, drv.accept_log = nullptr;
    // raw stream: one sweep consumes at most 2*ceil(3*nw*N/2) (Box-Muller) + nw*N (accept tests) outputs
    const unsigned long long gcap = 3ull * cap * N;
    const unsigned long long bcap = 2 * ((gcap + 1) / 2) + (unsigned long long)cap * N;
    set_backlog();
    unsigned long long ring = 1;
    while (ring < 3 * bcap + 2 * 624)
      ring <<= 1;
    A(rng_state, 624);
    A(rng_ring, ring);
    A(rng_cnt, 4);
    A(rng_flags, (size_t)cap + 1);
    rng.state = rng_state.p, rng.ring = rng_ring.p, rng.gen = rng_cnt.p, rng.pos = rng_cnt.p + 1;
    rng.sweep = rng_flags.p + cap, rng.flags = rng_flags.p;
    rng.ring_mask = (unsigned)(ring - 1);
    hd_abort();
    setup_segment_kernel(p->sweep_kernel);
    // oversubscribed two-kernel sweeps: walker indices by dispatch-order ticket (driver.cuh)
    {
      A(ticket, 1);
      int occ = 0, sms = 0, dev = 0;
      const int nmx     = std::max(det[0].n, det[1].n);
      const size_t smem = (size_t)(10 * nmx + k * (k + 1) + 5 * k) * sizeof(V);
      if (smem > 48 * 1024 && smem <= 200 * 1024)
        ensure_dynamic_smem(move_boundary_kernel<T, V>, 200 * 1024);
      QMCB_CUDA(cudaGetDevice(&dev));
      QMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      QMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, move_boundary_kernel<T, V>, MB_TPB, smem));
      const bool over = cap > occ * sms;
      drv.ticket      = (over || env_flag("QMCB_TICKET", 0)) ? ticket.p : nullptr;
      boundary_slots  = occ * sms;
    }
    mt19937_seed_kernel<<<1, 32, 0, st>>>(rng, p->seed);
    QMCB_LAUNCH_CHECK();
    mt19937_fill_kernel<<<1, 256, 0, st>>>(rng, 2 * sweep_backlog);
    QMCB_LAUNCH_CHECK();
    sync();
    vmc_ready = true;
  }

  // ---------------------------------------------------------------- per-kernel attribution of one sweep (measurement)
  struct ProfRec
  {
    int kind; // 0 walker-segment kernel, 1 boundary kernel, 2 spline gather, 3 Woodbury flush
    cudaEvent_t a, b;
  };
  std::vector<ProfRec>* prof = nullptr;
  void prof_begin(int kind)
  {
    if (!prof)
      return;
    ProfRec r;
    r.kind = kind;
    QMCB_CUDA(cudaEventCreate(&r.a));
    QMCB_CUDA(cudaEventCreate(&r.b));
    QMCB_CUDA(cudaEventRecord(r.a, st));
    prof->push_back(r);
  }
  void prof_end()
  {
    if (prof)
      QMCB_CUDA(cudaEventRecord(prof->back().b, st));
  }
  // out[0] sweep time (us), out[1 + kind] summed kernel time per kind, out[5 + kind] launches per kind
  void vmc_profile_sweep(double* out) override
  {
    flush_pending();
    if (!vmc_ready)
      throw std::runtime_error("qmcb_vmc_init has not been called");
    if (delay_count[0] != 0 || delay_count[1] != 0)
      twf_complete_updates();
    std::vector<ProfRec> recs;
    cudaEvent_t t0, t1;
    QMCB_CUDA(cudaEventCreate(&t0));
    QMCB_CUDA(cudaEventCreate(&t1));
    sync();
    prof = &recs;
    QMCB_CUDA(cudaEventRecord(t0, st));
    try
    {
      enqueue_sweep(false);
    }
    catch (...)
    {
      prof = nullptr;
      throw;
    }
    prof = nullptr;
    QMCB_CUDA(cudaEventRecord(t1, st));
    sync();
    for (int i = 0; i < 9; ++i)
      out[i] = 0.0;
    float ms = 0.f;
    QMCB_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    out[0] = 1e3 * ms;
    for (auto& r : recs)
    {
      QMCB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      out[1 + r.kind] += 1e3 * ms;
      out[5 + r.kind] += 1.0;
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
  }

  // ---------------------------------------------------------------- persistent walker-segment kernel (segment.cuh)
  void release_segment_slots()
  {
    if (seg_reserved_sms > 0)
      SegBudget::release(device, seg_reserved_sms);
    seg_reserved_sms = 0;
    fused            = false;
    hd.enabled       = false;
    hd.tried         = false;
  }
  template<int CPT, bool HD>
  int segment_occupancy(int spin, size_t& smem)
  {
    if constexpr (std::is_same<T, V>::value)
    {
      const SegLayout L = seg_layout<T, CPT>(det[spin].n, k, N, jas.has_j1 ? jas.nions : 0);
      smem              = L.total;
      if (smem > 227 * 1024)
        return 0;
      ensure_dynamic_smem(walker_segment_kernel<T, CPT, HD>, smem);
      int occ = 0;
      QMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, walker_segment_kernel<T, CPT, HD>, SEG_TPB, smem));
      return occ;
    }
    else
      return 0;
  }
  // can this crowd run the walker-segment kernel (device-driven, and host-driven when `host_mode`), with every CTA
  // resident?  Reserves the SM share once per crowd.  Returns "" or the reason why not.
  std::string acquire_segment_kernel(bool host_mode)
  {
    std::string why;
    if (!std::is_same<T, V>::value)
      why = "complex orbitals";
    else if (nmax > 2 * SEG_BOXW)
      why = "more than 384 orbitals per spin";
    else if (N + (jas.has_j1 ? jas.nions : 0) + 256 > 65535)
      why = "too many particles for the 16-bit cutoff lists";
    int occ = 1 << 30;
    if (why.empty())
      for (int spin = 0; spin < 2; ++spin)
        if (nel[spin] > 0)
        {
          size_t smem = 0;
          int o       = nel[spin] <= SEG_BOXW ? segment_occupancy<1, false>(spin, smem) : segment_occupancy<2, false>(spin, smem);
          if (host_mode)
            o = std::min(o, nel[spin] <= SEG_BOXW ? segment_occupancy<1, true>(spin, smem) : segment_occupancy<2, true>(spin, smem));
          occ = std::min(occ, o);
        }
    if (why.empty() && occ <= 0)
      why = "the walker's working set does not fit in shared memory";
    if (why.empty() && seg_reserved_sms <= 0)
    {
      int sms = 0;
      QMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      const double want = (double)cap / (double)occ;
      if (SegBudget::reserve(device, want, sms))
        seg_reserved_sms = want;
      else
        why = "not every walker's CTA can be resident at once (" + std::to_string(cap) + " walkers, " + std::to_string(occ) +
            " CTAs per SM, " + std::to_string(sms) + " SMs shared with the other crowds of this device)";
    }
    return why;
  }
  // mode: qmcb_vmc_params.sweep_kernel (0 automatic, 1 two-kernel path, 2 segment kernel required)
  void setup_segment_kernel(int mode)
  {
    release_segment_slots();
    {
      const char* e = std::getenv("QMCB_SWEEP_KERNEL");
      if (e && mode == 0)
        mode = std::atoi(e);
    }
    if (mode == 1)
      return;
    const std::string why = acquire_segment_kernel(false);
    if (!why.empty())
    {
      if (mode == 2)
        throw std::runtime_error("qmcb_vmc_init: the walker-segment kernel was requested but is not available: " + why);
      return;
    }
    fused = true;
    A(seg_flags, (size_t)N * cap);
    A(seg_tot_val, (size_t)N + 1);
    A(seg_tot_tag, (size_t)N + 1);
    A(seg_btot, (size_t)N * 32);
    segrng.flags = seg_flags.p, segrng.tot_val = seg_tot_val.p, segrng.tot_tag = seg_tot_tag.p, segrng.stride = cap;
    segrng.btot = seg_btot.p;
  }
  int vmc_sweep_kernel() const override { return vmc_ready ? (fused ? 2 : 1) : 0; }
  int host_kernel() const override { return hd.tried ? (hd.enabled ? 2 : 1) : 0; }
  // moves e0 .. e0 + nm - 1 of determinant `spin` in ONE launch (no flush inside: delay_count + nm <= k)
  void launch_segment(int spin, int e0, int nm)
  {
    if constexpr (std::is_same<T, V>::value)
    {
      const DetDev<V>& D = det[spin];
      if (delay_count[spin] + nm > k || e0 + nm > D.n)
        throw std::runtime_error("launch_segment: the segment crosses a flush or the end of the determinant");
      const SplineDev<T>& S  = *static_cast<const SplineDev<T>*>(spo[spin]->dev_desc());
      const CUtensorMap& tm = *static_cast<const CUtensorMap*>(spo[spin]->seg_tensor_map());
      const SegHost<T> noH{};
      if (D.n <= SEG_BOXW)
      {
        const SegLayout L = seg_layout<T, 1>(D.n, k, N, jas.has_j1 ? jas.nions : 0);
        walker_segment_kernel<T, 1, false><<<nw, SEG_TPB, L.total, st>>>(tm, S, drv, jas, rng, segrng, D, first[spin] + e0, e0, nm,
                                                                           delay_count[spin], noH);
      }
      else
      {
        const SegLayout L = seg_layout<T, 2>(D.n, k, N, jas.has_j1 ? jas.nions : 0);
        walker_segment_kernel<T, 2, false><<<nw, SEG_TPB, L.total, st>>>(tm, S, drv, jas, rng, segrng, D, first[spin] + e0, e0, nm,
                                                                           delay_count[spin], noH);
      }
      QMCB_LAUNCH_CHECK();
      delay_count[spin] += nm;
      invrow_id[spin] = -1;
    }
    else
      throw std::runtime_error("launch_segment: real orbitals only");
  }

  // ---------------------------------------------------------------- host-driven persistent segment (mailboxes, segment.cuh)
  // The per-electron calls of the reference's driver loop (mw_evalGrad, mw_makeMove, mw_calcRatioGrad,
  // mw_accept_rejectMove) normally cost one or two kernel launches and a stream round trip each.  When the calls arrive in
  // the loop's order, ONE resident kernel serves a whole segment (the moves up to the next flush) and every call becomes
  // a mailbox exchange in pinned host memory.  Any other call aborts the kernel and continues on the launch-per-call path.
  struct HostDrive
  {
    bool enabled = false, tried = false, active = false, grad_seen = false;
    PinBuf<unsigned char> pin;
    DevBuf<unsigned char> dev;
    SegHost<T> H{};
    unsigned seq = 0;   // last sequence number handed out
    unsigned seq0 = 0;  // first exchange of the active launch
    int iat0 = 0, iat = -1, iat_end = -1, spin = 0, c0 = 0;
    int stage = 0; // 0: gradient posted by the kernel / displacement expected, 1: ratio expected, 2: accept flags expected
  } hd;
  void hd_setup()
  {
    hd.tried = true;
    {
      const char* e = std::getenv("QMCB_HOST_KERNEL");
      if (e && std::string(e) == "launch")
        return;
    }
    if (!acquire_segment_kernel(true).empty())
      return;
    // pinned block of 16-byte chunks (segment.cuh, SegHost): h_out1 [cap][3] | h_out2 [cap][4] | h_in1 [cap] (float) or
    // [cap][3] (double) | h_in2 [ceil(cap / 4)] | h_cmd [4] u32
    auto up16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t o = 0;
    const size_t o_out1 = o;
    o += 16 * 3 * (size_t)cap;
    const size_t o_out2 = o;
    o += 16 * 4 * (size_t)cap;
    const size_t o_in1 = o;
    o += 16 * (sizeof(T) == 4 ? 1 : 3) * (size_t)cap;
    const size_t o_in2 = o;
    o += 16 * (((size_t)cap + 3) / 4);
    const size_t o_cmd = o;
    o += 64;
    hd.pin.alloc(o);
    std::memset(hd.pin.p, 0, o);
    hd.H.h_out1 = reinterpret_cast<uint4*>(hd.pin.p + o_out1);
    hd.H.h_out2 = reinterpret_cast<uint4*>(hd.pin.p + o_out2);
    hd.H.h_in1  = reinterpret_cast<const uint4*>(hd.pin.p + o_in1);
    hd.H.h_in2  = reinterpret_cast<const uint4*>(hd.pin.p + o_in2);
    hd.H.h_cmd  = reinterpret_cast<volatile unsigned*>(hd.pin.p + o_cmd);
    // device block: d_displ [3 cap] T | d_cmd [4] u32 | d_acc [cap]
    size_t d = 0;
    const size_t d_dis = d;
    d += up16(3 * (size_t)cap * sizeof(T));
    const size_t d_cmd = d;
    d += 64;
    const size_t d_acc = d;
    d += up16((size_t)cap);
    hd.dev.alloc(d);
    dev_bytes += hd.dev.bytes();
    hd.H.d_displ = reinterpret_cast<T*>(hd.dev.p + d_dis);
    hd.H.d_cmd   = reinterpret_cast<volatile unsigned*>(hd.dev.p + d_cmd);
    hd.H.d_acc   = hd.dev.p + d_acc;
    hd.enabled   = true;
  }
  // start a resident kernel for the moves iat .. end of the segment; false: not possible, use the launch-per-call path
  bool hd_begin(int iat)
  {
    if constexpr (std::is_same<T, V>::value)
    {
      if (!hd.tried)
        hd_setup();
      if (!hd.enabled)
        return false;
      join_jastrow();
      apply_pending(-1, nullptr); // an accept deferred by the launch-per-call path
      const int spin = spin_of(iat), row = iat - first[spin];
      const int nm   = std::min(k - delay_count[spin], nel[spin] - row);
      if (nm <= 0)
        return false;
      const DetDev<V>& D   = det[spin];
      const SplineDev<T>& S = *static_cast<const SplineDev<T>*>(spo[spin]->dev_desc());
      const CUtensorMap& tm = *static_cast<const CUtensorMap*>(spo[spin]->seg_tensor_map());
      hd.seq0   = hd.seq + 1;
      hd.H.seq0 = hd.seq0;
      hd.seq += 2u * (unsigned)nm;
      const_cast<volatile unsigned*>(hd.H.h_cmd)[0] = 0u; // (no abort requested)
      const_cast<volatile unsigned*>(hd.H.h_cmd)[1] = 0u;
      const unsigned init[2] = {hd.seq0 - 1u, 0u};
      QMCB_CUDA(cudaMemcpyAsync(const_cast<unsigned*>(hd.H.d_cmd), init, sizeof(init), cudaMemcpyHostToDevice, st));
      drv_host.nw = nw;
      if (D.n <= SEG_BOXW)
      {
        const SegLayout L = seg_layout<T, 1>(D.n, k, N, jas.has_j1 ? jas.nions : 0);
        walker_segment_kernel<T, 1, true><<<nw, SEG_TPB, L.total, st>>>(tm, S, drv_host, jas, rng, segrng, D, iat, row, nm,
                                                                          delay_count[spin], hd.H);
      }
      else
      {
        const SegLayout L = seg_layout<T, 2>(D.n, k, N, jas.has_j1 ? jas.nions : 0);
        walker_segment_kernel<T, 2, true><<<nw, SEG_TPB, L.total, st>>>(tm, S, drv_host, jas, rng, segrng, D, iat, row, nm,
                                                                          delay_count[spin], hd.H);
      }
      QMCB_LAUNCH_CHECK();
      hd.active = true, hd.grad_seen = false;
      hd.iat0 = hd.iat = iat, hd.iat_end = iat + nm, hd.spin = spin, hd.c0 = delay_count[spin], hd.stage = 0;
      invrow_id[spin] = -1;
      return true;
    }
    else
    {
      hd.tried = true; // (complex orbitals: the per-electron calls are launches; host_kernel() reports 1)
      return false;
    }
  }
  unsigned hd_seq(int iat, int second) const { return hd.seq0 + 2u * (unsigned)(iat - hd.iat0) + (unsigned)second; }
  // one 16-byte chunk {lo, hi} of a mailbox written by ONE store, so that the kernel's 16-byte read sees the tag (in the
  // high half) only together with the payload
  static void store16(void* dst, unsigned long long lo, unsigned long long hi)
  {
#if defined(__x86_64__)
    _mm_store_si128(static_cast<__m128i*>(dst), _mm_set_epi64x((long long)hi, (long long)lo)); // (aligned: one 16-byte store)
#else
    volatile unsigned long long* d = static_cast<volatile unsigned long long*>(dst);
    d[0] = lo;
    std::atomic_thread_fence(std::memory_order_release);
    d[1] = hi; // (the tag sits in the half written last)
#endif
  }
  // wait until chunk `slot` of an outgoing mailbox carries the tag of exchange `seq`; returns its 8 payload bytes
  unsigned long long hd_wait_chunk(const uint4* slot, unsigned seq, unsigned long long& spins)
  {
    const volatile unsigned* w = reinterpret_cast<const volatile unsigned*>(slot);
    while (w[2] != seq)
    {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 0x3fffffull) == 0)
      {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady)
        {
          hd.active = false;
          QMCB_CUDA(e);
          throw std::runtime_error("the resident walker-segment kernel ended before answering (time-out inside the kernel?)");
        }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    // (the device wrote the chunk with one 16-byte store: tag visible => payload visible)
    return (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  }
  static T bits_to_T(unsigned long long b)
  {
    if constexpr (sizeof(T) == 4)
    {
      const unsigned u = (unsigned)b;
      float f;
      std::memcpy(&f, &u, 4);
      return (T)f;
    }
    else
    {
      double d;
      std::memcpy(&d, &b, 8);
      return (T)d;
    }
  }
  // leave the resident kernel (no-op when none is active): the completed moves stay applied, a proposed but undecided
  // move is dropped exactly as if its accept had never been called
  void hd_abort()
  {
    if (!hd.active)
      return;
    std::atomic_thread_fence(std::memory_order_release);
    const_cast<volatile unsigned*>(hd.H.h_cmd)[0] = SEG_ABORT;
    QMCB_CUDA(cudaStreamSynchronize(st));
    unsigned done[2] = {0, 0};
    QMCB_CUDA(cudaMemcpy(done, const_cast<unsigned*>(hd.H.d_cmd), sizeof(done), cudaMemcpyDeviceToHost));
    hd.active              = false;
    delay_count[hd.spin]   = hd.c0 + (int)done[1];
    invrow_id[hd.spin]     = -1;
    const_cast<volatile unsigned*>(hd.H.h_cmd)[0] = 0u;
  }
  // exchange 1, kernel -> host: the component-summed gradient of the prepared electron of every live walker
  void hd_read_grads(unsigned seq, double* grads)
  {
    unsigned long long spins = 0;
    for (int iw = 0; iw < nw; ++iw)
      for (int d = 0; d < 3; ++d)
      {
        const unsigned long long b = hd_wait_chunk(hd.H.h_out1 + 3 * (size_t)iw + d, seq, spins);
        if (grads)
          grads[3 * (size_t)iw + d] = (double)bits_to_T(b);
      }
  }
  // the four calls of the move loop; each returns false when the call has to take the launch-per-call path
  bool hd_eval_grad(int iat, double* grads)
  {
    if (!hd.active && !hd_begin(iat))
      return false;
    if (hd.stage != 0 || iat != hd.iat)
    {
      hd_abort();
      return false;
    }
    hd_read_grads(hd_seq(iat, 0), grads);
    hd.grad_seen = true;
    return true;
  }
  bool hd_make_move(int iat, const double* dsp)
  {
    if (!hd.active && !hd_begin(iat))
      return false;
    if (hd.stage != 0 || iat != hd.iat)
    {
      hd_abort();
      return false;
    }
    if (!hd.grad_seen)
      hd_read_grads(hd_seq(iat, 0), nullptr); // (every CTA must have consumed the previous reply before this one is written)
    {
      const unsigned seq = hd_seq(iat, 0);
      unsigned char* h   = reinterpret_cast<unsigned char*>(const_cast<uint4*>(hd.H.h_in1));
      for (int iw = 0; iw < nw; ++iw)
      {
        if constexpr (sizeof(T) == 4)
        {
          const float f[3] = {(float)dsp[3 * iw], (float)dsp[3 * iw + 1], (float)dsp[3 * iw + 2]};
          unsigned u[3];
          std::memcpy(u, f, 12);
          store16(h + 16 * (size_t)iw, (unsigned long long)u[0] | ((unsigned long long)u[1] << 32),
                  (unsigned long long)u[2] | ((unsigned long long)seq << 32));
        }
        else
          for (int d = 0; d < 3; ++d)
          {
            unsigned long long b;
            std::memcpy(&b, &dsp[3 * iw + d], 8);
            store16(h + 16 * (size_t)(3 * iw + d), b, (unsigned long long)seq);
          }
      }
    }
    hd.stage      = 1;
    last_move_iat = iat;
    return true;
  }
  bool hd_calc_ratio_grad(int iat, double* ratios, double* grads)
  {
    if (!hd.active)
      return false;
    if (hd.stage != 1 || iat != hd.iat)
    {
      hd_abort();
      return false;
    }
    {
      const unsigned seq       = hd_seq(iat, 1);
      unsigned long long spins = 0;
      for (int iw = 0; iw < nw; ++iw)
      {
        const unsigned long long rb = hd_wait_chunk(hd.H.h_out2 + 4 * (size_t)iw, seq, spins);
        std::memcpy(&ratios[iw], &rb, 8);
        for (int d = 0; d < 3; ++d)
          grads[3 * (size_t)iw + d] = (double)bits_to_T(hd_wait_chunk(hd.H.h_out2 + 4 * (size_t)iw + 1 + d, seq, spins));
      }
    }
    hd.stage = 2;
    return true;
  }
  bool hd_accept_reject(int iat, const uint8_t* acc, int safe_to_delay)
  {
    if (!hd.active)
      return false;
    if (hd.stage != 2 || iat != hd.iat)
    {
      hd_abort();
      return false;
    }
    {
      // one word per walker: (sequence number << 1) | accepted, four walkers per 16-byte chunk
      const unsigned tag = (hd_seq(iat, 1) & 0x7fffffffu) << 1;
      unsigned char* h   = reinterpret_cast<unsigned char*>(const_cast<uint4*>(hd.H.h_in2));
      for (int e = 0; 4 * e < nw; ++e)
      {
        unsigned w4[4] = {0u, 0u, 0u, 0u};
        for (int q = 0; q < 4 && 4 * e + q < nw; ++q)
          w4[q] = tag | (acc[4 * e + q] ? 1u : 0u);
        store16(h + 16 * (size_t)e, (unsigned long long)w4[0] | ((unsigned long long)w4[1] << 32),
                (unsigned long long)w4[2] | ((unsigned long long)w4[3] << 32));
      }
    }
    hd.iat++;
    hd.stage     = 0;
    hd.grad_seen = false;
    if (hd.iat == hd.iat_end)
    {
      // the kernel applies this accept and ends by itself; whatever is launched next queues behind it on the stream
      hd.active            = false;
      delay_count[hd.spin] = hd.c0 + (hd.iat_end - hd.iat0);
      invrow_id[hd.spin]   = -1;
      if (delay_count[hd.spin] == k || !safe_to_delay)
        launch_flush(hd.spin, true);
    }
    else if (!safe_to_delay)
    {
      hd_abort();
      launch_flush(hd.spin);
    }
    return true;
  }

  // one launch of move_boundary_kernel: accept of electron iat_prev (or -1) and row preparation of iat_next (or -1).
  // ext_flags / twf_grads_out select the host-driven mode (Metropolis test done by the caller).
  void launch_boundary(const DriverDev<T>& dr, int iat_prev, int iat_next, const unsigned char* ext_flags, V* twf_grads_out)
  {
    int igp = 0, rp = 0, ign = 0, rn = 0;
    if (iat_prev >= 0)
    {
      igp = spin_of(iat_prev);
      rp  = iat_prev - first[igp];
    }
    if (iat_next >= 0)
    {
      ign = spin_of(iat_next);
      rn  = iat_next - first[ign];
    }
    if (iat_prev >= 0 && iat_next >= 0 && igp != ign)
      throw std::runtime_error("launch_boundary: both parts must belong to one determinant");
    const int cp = iat_prev >= 0 ? delay_count[igp] : 0;
    int cn       = iat_next >= 0 ? delay_count[ign] : 0;
    if (iat_prev >= 0 && iat_next >= 0)
      cn = cp + 1; // the slot appended by part 1 of this very launch
    const int nmx     = std::max(iat_prev >= 0 ? det[igp].n : 0, iat_next >= 0 ? det[ign].n : 0);
    const size_t smem = (size_t)(10 * nmx + k * (k + 1) + 5 * k) * sizeof(V);
    if (smem > 200 * 1024)
      throw std::runtime_error("determinant too wide for the boundary kernel's shared-memory staging");
    if (smem > 48 * 1024)
      ensure_dynamic_smem(move_boundary_kernel<T, V>, 200 * 1024);
    launch_kernel(move_boundary_kernel<T, V>, dim3(nw), dim3(MB_TPB), smem, st, (g_pdl_mode & 3) >= 2, dr, jas, rng, det[igp],
                  iat_prev, rp, cp, rg.p, rg_nparts, phi_vgl.p, det[ign], iat_next, rn, cn, det_grads.p, ext_flags,
                  twf_grads_out);
    QMCB_LAUNCH_CHECK();
    if (iat_prev >= 0)
    {
      delay_count[igp]++;
      invrow_id[igp] = -1;
    }
    if (iat_next >= 0)
      invrow_id[ign] = rn;
  }

  // ---- host-driven mode: the accept of an electron is deferred until the next call tells what follows it, so that
  // accept(iat) + evalGrad(iat+1) -- two consecutive calls of the reference's driver loop -- become ONE launch
  int pending_iat = -1;
  bool pending_value_only = false;
  DriverDev<T> drv_host{};
  void apply_pending(int iat_next, V* twf_grads_out)
  {
    if (pending_iat < 0)
    {
      if (iat_next >= 0)
        launch_boundary(drv_host, -1, iat_next, accepted.p, twf_grads_out);
      return;
    }
    const unsigned char* flags = pending_flags ? pending_flags : accepted.p;
    const int prev = pending_iat, ig = spin_of(prev);
    pending_iat    = -1;
    DriverDev<T> drv_host = this->drv_host; // (per-launch copy: the value-only flag belongs to this accept)
    drv_host.value_only   = pending_value_only ? 1 : 0;
    pending_value_only    = false;
    const bool split = iat_next < 0 || spin_of(iat_next) != ig || delay_count[ig] + 1 == k;
    if (!split)
    {
      launch_boundary(drv_host, prev, iat_next, flags, twf_grads_out);
      return;
    }
    launch_boundary(drv_host, prev, -1, flags, nullptr);
    if (delay_count[ig] == k)
      launch_flush(ig);
    if (iat_next >= 0)
      launch_boundary(drv_host, -1, iat_next, accepted.p, twf_grads_out);
  }
  void flush_pending()
  {
    hd_abort();
    join_jastrow();
    apply_pending(-1, nullptr);
  }

  // enqueue one full sweep on `st` (st2 carries the RNG top-up and the Jastrow branch)
  void enqueue_sweep(bool log_accept)
  {
    drv.accept_log = log_accept ? accept_log.p : nullptr;
    if (drv.dmc)
    {
      QMCB_CUDA(cudaMemsetAsync(rr_acc.p, 0, rr_acc.bytes(), st));
      QMCB_CUDA(cudaMemsetAsync(rr_prop.p, 0, rr_prop.bytes(), st));
    }
    // fork: top the raw stream up for the NEXT sweep on a side stream while this sweep runs.  The fill kernel may read a
    // stale (smaller) consumption counter: it then generates less, but the invariant "at least one sweep's worth is
    // available at sweep start" holds because it tops up to TWO sweeps' worth and a sweep consumes at most one.
    QMCB_CUDA(cudaEventRecord(ev_rng, st));
    QMCB_CUDA(cudaStreamWaitEvent(st3, ev_rng, 0));
    mt19937_fill_kernel<<<1, 256, 0, st3>>>(rng, 2 * sweep_backlog);
    QMCB_LAUNCH_CHECK();
    QMCB_CUDA(cudaEventRecord(ev_rng_done, st3));
    const unsigned long long gcount = 3ull * nw * N;
    gauss_kernel<T><<<(unsigned)(((gcount + 1) / 2 + 255) / 256), 256, 0, st>>>(rng, deltas.p, gcount, N & 1);
    QMCB_LAUNCH_CHECK();
    rng_advance_kernel<<<1, 32, 0, st>>>(rng, 2 * ((gcount + 1) / 2), N & 1);
    QMCB_LAUNCH_CHECK();
    if (fused)
    {
      // one launch per segment of <= delay_rank moves, then the flush: 2 launches per k moves instead of 2 per move
      for (int spin = 0; spin < 2; ++spin)
        for (int e0 = 0; e0 < nel[spin]; e0 += k)
        {
          prof_begin(0);
          launch_segment(spin, e0, std::min(k, nel[spin] - e0));
          prof_end();
          prof_begin(3);
          launch_flush(spin, true);
          prof_end();
        }
      twf_complete_updates();
      QMCB_CUDA(cudaStreamWaitEvent(st, ev_rng_done, 0));
      return;
    }
    // per electron: [accept(iat-1) + prepare/propose(iat)] -> {spline gather || Jastrow rows}; the boundary kernel is
    // split in two around a Woodbury flush
    auto boundary = [&](int iat_prev, int iat_next) {
      prof_begin(1);
      launch_boundary(drv, iat_prev, iat_next, nullptr, nullptr);
      prof_end();
    };
    boundary(-1, 0);
    for (int iat = 0; iat < N; ++iat)
    {
      const int ig = spin_of(iat);
      const bool jast = (jas.has_j2 || jas.has_j1) && !drv.fuse_jastrow;
      if (jast)
      {
        QMCB_CUDA(cudaEventRecord(ev_fork, st));
        QMCB_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
        jastrow_move_kernel<T, false><<<nw, JAS_TPB, 0, st2>>>(jas, iat);
        QMCB_LAUNCH_CHECK();
        QMCB_CUDA(cudaEventRecord(ev_join, st2));
      }
      prof_begin(2);
      launch_spline(ig, MODE_VGL, invRow[ig].p, det[ig].n, phi_vgl.p, rg.p, st);
      prof_end();
      if (jast)
        QMCB_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
      const int next = iat + 1 < N ? iat + 1 : -1;
      const bool flush_after = (delay_count[ig] + 1 == k) || (next >= 0 && spin_of(next) != ig);
      if (flush_after || next < 0)
      {
        boundary(iat, -1);
        prof_begin(3);
        launch_flush(ig);
        prof_end();
        if (next >= 0)
          boundary(-1, next);
      }
      else
        boundary(iat, next);
    }
    twf_complete_updates();
    // RNG top-up branch joins here
    QMCB_CUDA(cudaStreamWaitEvent(st, ev_rng_done, 0));
  }

  void vmc_sweep_async() override
  {
    flush_pending();
    if (!vmc_ready)
      throw std::runtime_error("qmcb_vmc_init has not been called");
    launch_sweep(false);
  }

  void launch_sweep(bool log_accept)
  {
    if (delay_count[0] != 0 || delay_count[1] != 0)
      twf_complete_updates();
    refresh_stale_gl();
    if (!use_graph)
    {
      enqueue_sweep(log_accept);
      return;
    }
    if (!graph_exec || graph_logs != log_accept)
    {
      if (graph_exec)
      {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
      }
      sync();
      const unsigned long long before = g_launch_count.load();
      cudaGraph_t graph;
      QMCB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      enqueue_sweep(log_accept);
      QMCB_CUDA(cudaStreamEndCapture(st, &graph));
      QMCB_CUDA(cudaGraphInstantiate(&graph_exec, graph, 0));
      QMCB_CUDA(cudaGraphDestroy(graph));
      graph_logs       = log_accept;
      launches_per_sweep = g_launch_count.load() - before;
      g_launch_count.store(before); // captured, not yet executed
    }
    QMCB_CUDA(cudaGraphLaunch(graph_exec, st));
    g_launch_count.fetch_add(launches_per_sweep);
  }
  unsigned long long launches_per_sweep = 0;

  void vmc_sweep(int nsteps, uint8_t* log_host) override
  {
    QMCB_NVTX("VMCBatched::advanceWalkers");
    flush_pending();
    if (!vmc_ready)
      throw std::runtime_error("qmcb_vmc_init has not been called");
    for (int s = 0; s < nsteps; ++s)
    {
      launch_sweep(log_host != nullptr);
      if (log_host)
      {
        QMCB_CUDA(cudaMemcpyAsync(log_host + (size_t)s * N * nw, accept_log.p, (size_t)N * nw, cudaMemcpyDeviceToHost, st));
        sync();
      }
    }
    sync();
    check_device_errors("VMCBatched::advanceWalkers (device-resident sweep)");
  }
  void vmc_counts(long long* na, long long* nr) override
  {
    QMCB_CUDA(cudaMemcpyAsync(na, n_acc.p, nw * sizeof(long long), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(nr, n_rej.p, nw * sizeof(long long), cudaMemcpyDeviceToHost, st));
    sync();
  }
  void dmc_get_rr(double* a, double* p) override
  {
    if (!vmc_ready || !drv.dmc)
      throw std::runtime_error("qmcb_dmc_get_rr: the driver was not initialised with dmc = 1");
    std::vector<T> ha(nw), hp(nw);
    QMCB_CUDA(cudaMemcpyAsync(ha.data(), rr_acc.p, nw * sizeof(T), cudaMemcpyDeviceToHost, st));
    QMCB_CUDA(cudaMemcpyAsync(hp.data(), rr_prop.p, nw * sizeof(T), cudaMemcpyDeviceToHost, st));
    sync();
    for (int i = 0; i < nw; ++i)
    {
      a[i] = (double)ha[i];
      p[i] = (double)hp[i];
    }
  }

  // ---------------------------------------------------------------- walker state (branching / load balancing)
  // every per-walker array is walker-major and contiguous: a walker's state is a list of (base, bytes per walker) segments
  struct Segment
  {
    unsigned char* base;
    size_t bytes;
  };
  std::vector<Segment> segments()
  {
    std::vector<Segment> sg;
    auto add = [&](void* p, size_t bytes) {
      if (p && bytes)
        sg.push_back({static_cast<unsigned char*>(p), bytes});
    };
    add(rsoa.p, 3 * npad * sizeof(T));
    for (int s2 = 0; s2 < 2; ++s2)
    {
      const size_t n = (size_t)nel[s2];
      add(Ainv[s2].p, n * lda[s2] * sizeof(V));
      add(GL[s2].p, n * 4 * n * sizeof(V));
      add(logdet[s2].p, 2 * sizeof(double));
    }
    if (jas.has_j2)
    {
      add(Uat.p, npad * sizeof(T));
      add(dUat.p, 3 * npad * sizeof(T));
      add(d2Uat.p, npad * sizeof(T));
      add(j2_log.p, sizeof(double));
    }
    if (jas.has_j1)
    {
      add(Vat.p, (size_t)N * sizeof(T));
      add(Grad1.p, 3 * (size_t)N * sizeof(T));
      add(Lap1.p, (size_t)N * sizeof(T));
      add(j1_log.p, sizeof(double));
    }
    return sg;
  }
  size_t walker_bytes() const override
  {
    size_t b = 0;
    for (auto& s2 : const_cast<Crowd*>(this)->segments())
      b += (s2.bytes + 15) & ~size_t(15);
    return b;
  }
  void set_backlog()
  {
    const unsigned long long gcount = 3ull * nw * N;
    sweep_backlog                   = 2 * ((gcount + 1) / 2) + (unsigned long long)nw * N;
  }
  // DMC population change: the buffers keep their capacity, every kernel and stride follows the live count
  void set_num_walkers(int n_active) override
  {
    if (n_active < 1 || n_active > cap)
      throw std::runtime_error("qmcb_crowd_set_num_walkers: count must be in [1, capacity]");
    settle();
    sync();
    nw = n_active;
    det[0].nw = det[1].nw = nw;
    jas.nw = nw, drv_host.nw = nw;
    if (vmc_ready)
    {
      drv.nw = nw;
      set_backlog();
    }
    if (graph_exec)
    {
      cudaGraphExecDestroy(graph_exec);
      graph_exec = nullptr;
    }
  }
  int num_walkers() const override { return nw; }
  int capacity() const override { return cap; }
  void check_walker(int iw) const
  {
    if (iw < 0 || iw >= cap)
      throw std::runtime_error("walker index out of range");
  }
  void settle()
  {
    flush_pending();
    refresh_stale_gl();
    if (delay_count[0] != 0 || delay_count[1] != 0)
      twf_complete_updates();
    invrow_id[0] = invrow_id[1] = -1;
  }
  void pack_walker(int iw, void* dev_buf) override
  {
    QMCB_NVTX("DiracDeterminantBatched::buffer");
    check_walker(iw);
    settle();
    unsigned char* out = static_cast<unsigned char*>(dev_buf);
    for (auto& s2 : segments())
    {
      QMCB_CUDA(cudaMemcpyAsync(out, s2.base + (size_t)iw * s2.bytes, s2.bytes, cudaMemcpyDeviceToDevice, st));
      out += (s2.bytes + 15) & ~size_t(15);
    }
    sync();
  }
  void unpack_walker(int iw, const void* dev_buf) override
  {
    check_walker(iw);
    settle();
    const unsigned char* in = static_cast<const unsigned char*>(dev_buf);
    for (auto& s2 : segments())
    {
      QMCB_CUDA(cudaMemcpyAsync(s2.base + (size_t)iw * s2.bytes, in, s2.bytes, cudaMemcpyDeviceToDevice, st));
      in += (s2.bytes + 15) & ~size_t(15);
    }
    sync();
  }
  void copy_walker(int src, int dst) override
  {
    check_walker(src);
    check_walker(dst);
    if (src == dst)
      return;
    settle();
    for (auto& s2 : segments())
      QMCB_CUDA(cudaMemcpyAsync(s2.base + (size_t)dst * s2.bytes, s2.base + (size_t)src * s2.bytes, s2.bytes,
                                cudaMemcpyDeviceToDevice, st));
    sync();
  }
};

CrowdBase* make_crowd(const qmcb_system* sys, int nw)
{
  // complex orbitals (SplineC2C tables) select the complex determinant engine, as QMC_COMPLEX does in the reference
  const bool cplx = sys->spo[0] && sys->spo[0]->impl && sys->spo[0]->impl->kind == QMCB_C2C;
  if (sys->precision == QMCB_MIXED)
    return cplx ? static_cast<CrowdBase*>(new Crowd<float, cx<float>>(sys, nw)) : new Crowd<float, float>(sys, nw);
  if (sys->precision == QMCB_FULL)
    return cplx ? static_cast<CrowdBase*>(new Crowd<double, cx<double>>(sys, nw)) : new Crowd<double, double>(sys, nw);
  throw std::runtime_error("unknown precision code");
}

} // namespace qmcb
