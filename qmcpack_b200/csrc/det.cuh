// qmcpack_b200/csrc/det.cuh -- Slater-determinant delayed-update engine kernels (sm_100a).
//
// Replaces (reference paths under /root/reference/src/QMCWaveFunctions):
//   DelayedUpdateBatched::mw_prepareInvRow  Fermion/DelayedUpdateBatched.h:174-235   copy_batched + 3 gemv_batched
//   DelayedUpdateBatched::mw_evalGrad       :354-400 + detail/CUDA/matrix_update_helper.cu:182-295 calcGradients
//   DelayedUpdateBatched::mw_accept_rejectRow :542-670  copy + 2 gemv + ger + add_delay_list_save_sigma_VGL
//                                           (matrix_update_helper.cu:297-505; pseudo-accept of rejected walkers)
//   DelayedUpdateBatched::mw_updateInvMat   :675-738  3 gemm_batched + applyW_batched
// Arithmetic spec: SURVEY Appendix A.5 / oracle/qmc_oracle.hpp DelayedUpdate (explicit row-major indices).
//
// Layout (one determinant of one crowd; everything contiguous over walkers, no pointer arrays):
//   Ainv [nw][n][lda]   psiMinv = (psiM^-1)^T, row = electron, col = orbital
//   GL   [nw][n][4][n]  rows of dpsiM (x,y,z as three contiguous vectors) and d2psiM
//   U,V  [nw][k][n]     new orbital rows / saved (stale) inverse rows of the delayed electrons
//   Binv [nw][k][k]     inverse of the k x k Woodbury core, row-major
//   wvec [nw][k]        w = -Binv^T p left by the row preparation (the reference parks it in Binv row `count`)
//   list [nw][k]        delayed electron index, -1 for a pseudo-accepted (rejected) slot
// The reference issues 4 launches to prepare a row and 5 to accept one; here each is ONE kernel per crowd.
#pragma once
#include "common.cuh"

namespace qmcb
{
template<typename T>
struct DetDev
{
  int n, lda, k, nw;
  T* Ainv;
  T* GL;
  T* U;
  T* V;
  T* Binv;
  T* wvec;
  int* list;
  T* invRow;      // [nw][n]
  T* tempMat;     // [nw][n][k]
  T* Up;          // [nw][k][n]
  double* logdet; // [nw][2]
  unsigned* err;  // crowd-wide error bits set by the kernels, read by the host at its next synchronisation:
                  // 1 = a one-particle ratio was NaN (NaNguard::checkOneParticleRatio, TrialWaveFunction.cpp:473,508,549)
                  // 2 = a move was accepted with a zero determinant ratio (DiracDeterminantBatched.cpp:494-500)
};
constexpr unsigned QMCB_ERR_NAN_RATIO = 1u, QMCB_ERR_ZERO_RATIO_ACCEPTED = 2u;

#ifdef __CUDACC__
constexpr int DET_TPB = 256;

// log_value += log(curRatio) as a complex logarithm (DiracDeterminantBatched.cpp:501): ld = {log|r|, arg r}
__device__ __forceinline__ void logdet_accumulate(double* ld, const double r)
{
  ld[0] += log(fabs(r));
  if (r < 0)
    ld[1] += 3.14159265358979323846;
}
__device__ __forceinline__ void logdet_accumulate(double* ld, const float r) { logdet_accumulate(ld, (double)r); }
template<typename T>
__device__ __forceinline__ void logdet_accumulate(double* ld, const cx<T>& r)
{
  const double re = (double)r.re, im = (double)r.im;
  ld[0] += 0.5 * log(re * re + im * im);
  ld[1] += atan2(im, re);
}

// invRow = Ainv[row] - V^T (Binv^T (U Ainv[row]))  and  grad_now = invRow . dpsiM[row]   (walker iw, thread group g)
// x[n], p[k], w[k] shared scratch; red >= 3*32.  On return every thread of the group holds grad_now in gout (if asked).
// x_ready: x[] already holds Ainv[row]; p_ready: p[0..p_ready) = U[a].x already computed by the caller
template<typename T>
__device__ __forceinline__ void det_prepare_row_body(const Group& g, const DetDev<T>& D, const int iw, const int row,
                                                     const int c, T* x, T* p, T* w, T* red, const bool want_grads,
                                                     T gout[3], const bool x_ready = false, const int p_ready = 0)
{
  const int n = D.n, k = D.k;
  const int lane = g.tid & 31, warp = g.tid >> 5, nwarp = g.n >> 5;
  if (!x_ready)
  {
    const T* arow = D.Ainv + ((size_t)iw * n + row) * D.lda;
    for (int j = g.tid; j < n; j += g.n)
      x[j] = arow[j];
    g.sync();
  }
  if (c > 0)
  {
    const T* U = D.U + (size_t)iw * k * n;
    const T* V = D.V + (size_t)iw * k * n;
    const T* B = D.Binv + (size_t)iw * k * k;
    for (int a = p_ready + warp; a < c; a += nwarp)
    {
      T s(0);
      for (int j = lane; j < n; j += 32)
        s += U[(size_t)a * n + j] * x[j];
      s = warp_sum(s);
      if (lane == 0)
        p[a] = s;
    }
    g.sync();
    if (g.tid < c)
    {
      T s(0);
      for (int a = 0; a < c; ++a)
        s += B[a * k + g.tid] * p[a];
      w[g.tid]                        = -s;
      D.wvec[(size_t)iw * k + g.tid] = -s;
    }
    g.sync();
    for (int j = g.tid; j < n; j += g.n)
    {
      T s(0);
      for (int a = 0; a < c; ++a)
        s += V[(size_t)a * n + j] * w[a];
      x[j] += s;
    }
    g.sync();
  }
  T* out = D.invRow + (size_t)iw * n;
  for (int j = g.tid; j < n; j += g.n)
    out[j] = x[j];
  if (want_grads)
  {
    const T* gl = D.GL + ((size_t)iw * n + row) * 4 * n;
    T acc[3]    = {T(0), T(0), T(0)};
    for (int j = g.tid; j < n; j += g.n)
    {
      const T xv = x[j];
      acc[0] += xv * gl[j];
      acc[1] += xv * gl[n + j];
      acc[2] += xv * gl[2 * n + j];
    }
    group_sum<T, 3>(g, acc, red);
    gout[0] = acc[0];
    gout[1] = acc[1];
    gout[2] = acc[2];
  }
}

// grads (optional) [nw][3].  One CTA per walker.  dynamic smem: (n + 2k) * sizeof(T)
template<typename T>
__global__ void __launch_bounds__(DET_TPB) det_prepare_row_kernel(const DetDev<T> D, const int row, const int c, T* grads)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* x = reinterpret_cast<T*>(smem_raw);
  T* p = x + D.n;
  T* w = p + D.k;
  __shared__ T red[3 * 32];
  T gout[3];
  det_prepare_row_body<T>(cta_group(), D, blockIdx.x, row, c, x, p, w, red, grads != nullptr, gout);
  if (grads && threadIdx.x < 3)
    grads[(size_t)blockIdx.x * 3 + threadIdx.x] = gout[threadIdx.x];
}

// ratio/grad of externally supplied orbital rows (FakeSPO-style tests): rg[iw][4] = invRow . {v, gx, gy, gz}
template<typename T>
__global__ void __launch_bounds__(DET_TPB) det_ratio_from_phi_kernel(const DetDev<T> D, const T* phi_vgl, T* rg)
{
  __shared__ T red[4 * 32];
  const int iw = blockIdx.x, tid = threadIdx.x, n = D.n;
  const size_t fs = (size_t)D.nw * n;
  const T* x      = D.invRow + (size_t)iw * n;
  const T* ph     = phi_vgl + (size_t)iw * n;
  T acc[4]        = {T(0), T(0), T(0), T(0)};
  for (int j = tid; j < n; j += DET_TPB)
  {
    const T xv = x[j];
    acc[0] += xv * ph[j];
    acc[1] += xv * ph[fs + j];
    acc[2] += xv * ph[2 * fs + j];
    acc[3] += xv * ph[3 * fs + j];
  }
  block_sum<T, 4>(acc, red);
  if (tid < 4)
    rg[(size_t)iw * 4 + tid] = acc[tid];
}

// accept / pseudo-accept of slot c for walker iw by thread group g.  phi[n], p[k], y[k] shared scratch.
// have_p: phi[] already holds the new orbital values and p[a] = -V[a].phi (a < c) was computed by the caller
template<typename T>
__device__ __forceinline__ void det_accept_body(const Group& g, const DetDev<T>& D, const int iw, const int row, const int c,
                                                const bool acc, const T ratio, const T* phi_vgl, T* phi, T* p, T* y,
                                                const bool have_p = false)
{
  const int n = D.n, k = D.k;
  const int lane = g.tid & 31, warp = g.tid >> 5, nwarp = g.n >> 5;
  T* U            = D.U + (size_t)iw * k * n;
  T* V            = D.V + (size_t)iw * k * n;
  T* B            = D.Binv + (size_t)iw * k * k;
  const T* arow   = D.Ainv + ((size_t)iw * n + row) * D.lda;
  const size_t fs = (size_t)D.nw * n;
  // V[c] = Ainv[row] (stale stored row) for every walker, DelayedUpdateBatched.h:646
  for (int j = g.tid; j < n; j += g.n)
    V[(size_t)c * n + j] = arow[j];
  if (acc && g.tid == 0 && ratio == T(0) && D.err)
    atomicOr(D.err, QMCB_ERR_ZERO_RATIO_ACCEPTED);
  if (acc)
  {
    const T* ph = phi_vgl + (size_t)iw * n;
    T* gl       = D.GL + ((size_t)iw * n + row) * 4 * n;
    for (int j = g.tid; j < n; j += g.n)
    {
      const T v = have_p ? phi[j] : ph[j];
      phi[j]    = v;
      U[(size_t)c * n + j] = v;
      gl[j]         = ph[fs + j];
      gl[n + j]     = ph[2 * fs + j];
      gl[2 * n + j] = ph[3 * fs + j];
      gl[3 * n + j] = ph[4 * fs + j];
    }
    g.sync();
    if (!have_p)
    {
      for (int a = warp; a < c; a += nwarp)
      {
        T s(0);
        for (int j = lane; j < n; j += 32)
          s += V[(size_t)a * n + j] * phi[j];
        s = warp_sum(s);
        if (lane == 0)
          p[a] = -s;
      }
      g.sync();
    }
    const T sigma = T(1) / ratio;
    const T* w    = D.wvec + (size_t)iw * k;
    if (g.tid < c)
    {
      T s(0);
      for (int b = 0; b < c; ++b)
        s += B[g.tid * k + b] * p[b];
      y[g.tid] = sigma * s;
    }
    g.sync();
    for (int e = g.tid; e < c * c; e += g.n)
    {
      const int a = e / c, b = e - a * c;
      B[a * k + b] += y[a] * w[b];
    }
    if (g.tid < c)
    {
      B[g.tid * k + c] = y[g.tid];
      B[c * k + g.tid] = sigma * w[g.tid];
    }
    if (g.tid == 0)
    {
      B[c * k + c]               = sigma;
      D.list[(size_t)iw * k + c] = row;
      logdet_accumulate(D.logdet + 2 * (size_t)iw, ratio);
    }
  }
  else
  {
    // pseudo-accept: detail/OMPTarget/AccelMatrixUpdateOMPTarget.hpp:139-160
    for (int j = g.tid; j < n; j += g.n)
      U[(size_t)c * n + j] = T(0);
    if (g.tid < c)
    {
      B[c * k + g.tid] = T(0);
      B[g.tid * k + c] = T(0);
    }
    if (g.tid == 0)
    {
      B[c * k + c]               = T(1);
      D.list[(size_t)iw * k + c] = -1;
    }
  }
}

// rg[iw][part][0] = partial determinant ratios of the proposed move.  dynamic smem: (n + 2k) * sizeof(T)
template<typename T>
__global__ void __launch_bounds__(DET_TPB)
    det_accept_kernel(const DetDev<T> D, const int row, const int c, const unsigned char* accepted, const T* rg,
                      const int rg_nparts, const T* phi_vgl)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* phi = reinterpret_cast<T*>(smem_raw);
  T* p   = phi + D.n;
  T* y   = p + D.k;
  const int iw = blockIdx.x;
  // determinant ratio = sum of the spline kernel's partial dots in index order
  T ratio(0);
  for (int q = 0; q < rg_nparts; ++q)
    ratio += rg[((size_t)iw * rg_nparts + q) * 4];
  det_accept_body<T>(cta_group(), D, iw, row, c, accepted[iw] != 0, ratio, phi_vgl, phi, p, y);
}

// ---- Woodbury flush, first (SIMT) form: three small batched GEMMs with the -1 fix-up fused into the first.
// C[M x N] (ldc) = alpha * A[M x K] (lda) * op(B) + beta * C;  op(B) = B[K x N] (ldb) or B^T with B[N x K] (ldb)
// 64x64 output tile per CTA, 256 threads, 4x4 micro-tile, K step 16.  grid = (N/64, M/64, batch)
template<typename T, bool TRANSB, bool FIXUP>
__global__ void __launch_bounds__(256)
    gemm_batched_kernel(int M, int N, int K, T alpha, const T* A, int lda, size_t strideA, const T* B, int ldb,
                        size_t strideB, T beta, T* C, int ldc, size_t strideC, const int* list, int list_stride)
{
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int b = blockIdx.z;
  A += (size_t)b * strideA;
  B += (size_t)b * strideB;
  C += (size_t)b * strideC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc[i][j] = T(0);
  for (int k0 = 0; k0 < K; k0 += BK)
  {
    // A tile: BM x BK, row-major source -> As[kk][m]
    for (int e = tid; e < BM * BK; e += 256)
    {
      const int m = e / BK, kk = e - m * BK;
      const int gm = m0 + m, gk = k0 + kk;
      As[kk][m] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : T(0);
    }
    if (TRANSB)
    {
      for (int e = tid; e < BN * BK; e += 256)
      {
        const int nn = e / BK, kk = e - nn * BK;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : T(0);
      }
    }
    else
    {
      for (int e = tid; e < BN * BK; e += 256)
      {
        const int kk = e / BN, nn = e - kk * BN;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < N && gk < K) ? B[(size_t)gk * ldb + gn] : T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk)
    {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }
  const int* lst = FIXUP ? list + (size_t)b * list_stride : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N)
      {
        T v = alpha * acc[i][j];
        if (beta != T(0))
          v += beta * C[(size_t)gm * ldc + gn];
        if (FIXUP && lst[gn] == gm) // applyW: tempMat[list[a]][a] -= 1 (skips -1 entries)
          v -= T(1);
        C[(size_t)gm * ldc + gn] = v;
      }
    }
}

// psiM scatter for the from-scratch path: phi_vgl [5][nw][n] of electron `e` -> transposed FP64 matrix
// AT[iw][j][e] = phi_j (so that the column-major LU of AT is the LU of psiM) and GL[iw][e][4][n]
template<typename T>
__global__ void det_scatter_row_kernel(const DetDev<T> D, const int e, const T* phi_vgl,
                                       typename value_traits<T>::dbl_t* AT)
{
  const int iw = blockIdx.y, n = D.n;
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  const size_t fs = (size_t)D.nw * n;
  const T* ph     = phi_vgl + (size_t)iw * n;
  AT[((size_t)iw * n + j) * n + e] = to_dbl(ph[j]);
  T* gl                            = D.GL + ((size_t)iw * n + e) * 4 * n;
  gl[j]                            = ph[fs + j];
  gl[n + j]                        = ph[2 * fs + j];
  gl[2 * n + j]                    = ph[3 * fs + j];
  gl[3 * n + j]                    = ph[4 * fs + j];
}

// gradient / Laplacian rows of electron `e` only (psiM is not touched): refresh of dpsiM, d2psiM after ratio-only moves
// (DiracDeterminantBatched::mw_evaluateGL with UpdateMode == ORB_PBYP_RATIO re-evaluates them, DiracDeterminantBatched.cpp:630-678)
template<typename T>
__global__ void det_scatter_gl_kernel(const DetDev<T> D, const int e, const T* phi_vgl)
{
  const int iw = blockIdx.y, n = D.n;
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  const size_t fs = (size_t)D.nw * n;
  const T* ph     = phi_vgl + (size_t)iw * n;
  T* gl           = D.GL + ((size_t)iw * n + e) * 4 * n;
  gl[j]           = ph[fs + j];
  gl[n + j]       = ph[2 * fs + j];
  gl[2 * n + j]   = ph[3 * fs + j];
  gl[3 * n + j]   = ph[4 * fs + j];
}

// log-determinant from the LU factors (DiracMatrix.h:100-107 / detail/CUDA/cuBLAS_LU.cu:61-110):
// sum_i log(complex(pivot[i]==i+1 ? diag : -diag)); LU is column-major [n][n] per walker
template<typename DT /* double or cx<double> */>
__global__ void det_logdet_kernel(const DT* LU, const int* piv, int n, double* logdet)
{
  __shared__ double red[2 * 32];
  const int iw = blockIdx.x, tid = threadIdx.x;
  double acc[2] = {0.0, 0.0};
  for (int i = tid; i < n; i += blockDim.x)
  {
    DT d = LU[((size_t)iw * n + i) * n + i];
    if (piv[(size_t)iw * n + i] != i + 1)
      d = -d;
    logdet_accumulate(acc, d);
  }
  block_sum<double, 2>(acc, red);
  if (tid == 0)
  {
    logdet[2 * (size_t)iw]     = acc[0];
    logdet[2 * (size_t)iw + 1] = acc[1];
  }
}

// inverse (column-major X^-1 == row-major (X^-1)^T) cast into Ainv [nw][n][lda]
template<typename T>
__global__ void det_store_inverse_kernel(const DetDev<T> D, const typename value_traits<T>::dbl_t* inv)
{
  const int iw = blockIdx.z, i = blockIdx.y, n = D.n;
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n)
    D.Ainv[((size_t)iw * n + i) * D.lda + j] = from_dbl<T>(inv[((size_t)iw * n + i) * n + j]);
}

// G, L of every electron of this determinant (DiracDeterminantBatched.cpp:594-604 computeGL):
// rv = invRow_i . dpsiM_i ; lap = invRow_i . d2psiM_i ; G += rv ; L += lap - rv.rv     out: Gd [nw][N][3], Ld [nw][N]
template<typename T, typename RT>
__global__ void __launch_bounds__(128) det_compute_gl_kernel(const DetDev<T> D, const int first, const int N, RT* Gd, RT* Ld)
{
  __shared__ T red[4 * 32];
  const int iw = blockIdx.y, i = blockIdx.x, tid = threadIdx.x, n = D.n;
  const T* inv = D.Ainv + ((size_t)iw * n + i) * D.lda;
  const T* gl  = D.GL + ((size_t)iw * n + i) * 4 * n;
  T acc[4]     = {T(0), T(0), T(0), T(0)};
  for (int j = tid; j < n; j += blockDim.x)
  {
    const T x = inv[j];
    acc[0] += x * gl[j];
    acc[1] += x * gl[n + j];
    acc[2] += x * gl[2 * n + j];
    acc[3] += x * gl[3 * n + j];
  }
  block_sum<T, 4>(acc, red);
  if (tid == 0)
  {
    const int iat = first + i;
    RT* g         = Gd + ((size_t)iw * N + iat) * 3;
    g[0] += (RT)acc[0];
    g[1] += (RT)acc[1];
    g[2] += (RT)acc[2];
    Ld[(size_t)iw * N + iat] += (RT)(acc[3] - (acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]));
  }
}
#endif

} // namespace qmcb
