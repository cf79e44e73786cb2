// qmcpack_b200/csrc/inverse.cuh -- batched FP64 inverse + log-determinant of the Slater matrices (sm_100a), the engine of
// DiracDeterminantBatched::mw_recompute / DiracMatrixInverterCUDA::mw_invertTranspose
// (Fermion/DiracDeterminantBatched.cpp:1122-1198, Fermion/DiracMatrixInverterCUDA.hpp:112-232, 306-369).
//
// The reference calls cublas{D,Z}getrfBatched + getriBatched (detail/CUDA/cuBLAS_LU.cu:61-210): LU with partial pivoting,
// the log-determinant from the pivots and the diagonal, then the inverse from the factors.  Those batched routines run one
// small CTA per matrix and are far from any roofline at n = 192 ... 1536.  Here: BLOCKED GAUSS-JORDAN with partial (row)
// pivoting, in place, as a sequence of panel steps over the whole batch:
//
//   panel kernel (one CTA per walker):  the n x b column panel goes to shared memory; b elimination steps choose the pivot
//       row of each column among the rows not used yet (largest modulus, lowest index on ties), and the in-place
//       Gauss-Jordan recurrence leaves  E + W  in the panel, where E holds the unit vectors of the pivot rows R and W
//       (n x b) is the whole effect of the b eliminations on the rest of the matrix:   X <- X + W X[R, :].
//   update kernel (grid over 64 x 64 tiles of every walker's matrix):  that rank-b update on the FP64 tensor pipe
//       (mma.sync.m8n8k4.f64 = DMMA, the fragment types of woodbury_dmma.cuh), X read and written once per panel step.
//
// Every step touches all n rows (the inverse accumulates in the columns already eliminated), so all panel steps cost
// the same: 2 n^2 b flops and one pass over the matrix; 2 n^3 in total, the count of getrf + getri.  At the end the
// stored matrix S is the inverse up to the row/column order of the pivots:  Xinv[c][j] = S[R[c]][Rinv[j]], and
// det X = sgn(R) prod(pivots) -- the finishing kernel applies the permutation while it casts into Ainv and adds the parity
// to the phase of the log-determinant.
//
// The matrix handed in is the row-major transpose of psiM (AT[j][e] = psiM[e][j]), so its row-major inverse IS
// psiMinv = (psiM^-1)^T, the layout the delayed-update engine works on (DiracMatrix.h:189-221).
#pragma once
#include "common.cuh"
#include "det.cuh"
#include "woodbury_dmma.cuh"

namespace qmcb
{
#ifdef __CUDACC__
namespace gj
{
constexpr int TPB  = 256;
constexpr int TILE = 64;

__device__ __forceinline__ double modulus2(const double v) { return v * v; }
__device__ __forceinline__ double modulus2(const cx<double>& v) { return v.re * v.re + v.im * v.im; }
__device__ __forceinline__ double one_over(const double v) { return 1.0 / v; }
__device__ __forceinline__ cx<double> one_over(const cx<double>& v) { return cx<double>(1.0) / v; }
__device__ __forceinline__ bool is_zero(const double v) { return v == 0.0; }
__device__ __forceinline__ bool is_zero(const cx<double>& v) { return v.re == 0.0 && v.im == 0.0; }

// two consecutive values of a 16-byte aligned shared-memory row
__device__ __forceinline__ void ld2(const double* p, double& a, double& b)
{
  const double2 v = *reinterpret_cast<const double2*>(p);
  a = v.x, b = v.y;
}
__device__ __forceinline__ void st2(double* p, const double a, const double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ void ld2(const cx<double>* p, cx<double>& a, cx<double>& b)
{
  const double2 v0 = reinterpret_cast<const double2*>(p)[0], v1 = reinterpret_cast<const double2*>(p)[1];
  a = cx<double>(v0.x, v0.y), b = cx<double>(v1.x, v1.y);
}
__device__ __forceinline__ void st2(cx<double>* p, const cx<double>& a, const cx<double>& b)
{
  reinterpret_cast<double2*>(p)[0] = make_double2(a.re, a.im);
  reinterpret_cast<double2*>(p)[1] = make_double2(b.re, b.im);
}

// shared-memory bytes of the panel kernel
// (rows of BW + 2 values: 16-byte aligned for the vector accesses of the elimination, and a stride of 4 (mod 32) banks
// per row, so that a quarter warp's 16-byte accesses to eight consecutive rows are conflict free)
template<typename V>
inline size_t panel_smem(int n, int BW)
{
  return (size_t)n * (BW + 2) * sizeof(V) + (size_t)BW * sizeof(V) + (size_t)(n + 15) / 16 * 16 + 64 * sizeof(int) +
         32 * sizeof(double) + 32 * sizeof(int);
}

// One panel step of one walker.  X [nw][n][n] row-major (in place); used [nw][n] row flags; R [nw][n] pivot row of every
// column; W [nw][n][BW], Xr [nw][BW][n] (BW = slots allocated per panel, b <= BW used); logdet [nw][2]; info [nw].
template<typename V, int BW>
__global__ void __launch_bounds__(TPB) gj_panel_kernel(V* __restrict__ X, const int n, const int j0, const int b,
                                                       unsigned char* __restrict__ used, int* __restrict__ R,
                                                       V* __restrict__ W, V* __restrict__ Xr, double* __restrict__ logdet,
                                                       int* __restrict__ info)
{
  extern __shared__ __align__(16) unsigned char gj_smem[];
  const int iw = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int ldp = BW + 2;
  V* Pn   = reinterpret_cast<V*>(gj_smem);
  V* prow = Pn + (size_t)n * ldp;
  unsigned char* us = reinterpret_cast<unsigned char*>(prow + BW);
  int* Rs           = reinterpret_cast<int*>(us + (size_t)(n + 15) / 16 * 16);
  double* rbest     = reinterpret_cast<double*>(Rs + 64);
  int* rbi          = reinterpret_cast<int*>(rbest + 32);
  V* Xw             = X + (size_t)iw * n * n;
  unsigned char* ug = used + (size_t)iw * n;

  // (a last, narrower panel is padded with zero columns: they stay zero through the eliminations)
  for (int e = tid; e < n * BW; e += TPB)
  {
    const int i = e / BW, k = e - i * BW;
    Pn[(size_t)i * ldp + k] = k < b ? Xw[(size_t)i * n + j0 + k] : V(0.0);
  }
  for (int i = tid; i < n; i += TPB)
    us[i] = j0 == 0 ? 0 : ug[i];
  double ld[2] = {0.0, 0.0};
  int bad      = 0;
  __syncthreads();

  for (int k = 0; k < b; ++k)
  {
    // ---- pivot: largest modulus among the rows not used yet; ties and NaNs resolve to the lowest index
    double best = -1.0;
    int bi      = n;
    for (int i = tid; i < n; i += TPB)
      if (!us[i])
      {
        const double a = modulus2(Pn[(size_t)i * ldp + k]);
        if (a > best)
          best = a, bi = i;
        else if (bi == n)
          bi = i; // (a NaN entry: still a row to fall back on)
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi    = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi))
        best = ob, bi = oi;
    }
    if (lane == 0)
      rbest[warp] = best, rbi[warp] = bi;
    __syncthreads();
    if (warp == 0)
    {
      best = lane < TPB / 32 ? rbest[lane] : -2.0;
      bi   = lane < TPB / 32 ? rbi[lane] : n;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi    = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi))
          best = ob, bi = oi;
      }
      if (lane == 0)
      {
        Rs[k]  = bi;
        us[bi] = 1;
      }
    }
    __syncthreads();
    const int p = Rs[k];
    const V pv  = Pn[(size_t)p * ldp + k];
    if (tid == 0)
    {
      logdet_accumulate(ld, pv);
      if (is_zero(pv) || !(modulus2(pv) == modulus2(pv)))
        bad = j0 + k + 1; // singular (or NaN) pivot: reported like the `info` of getrf
    }
    const V rp = one_over(pv);
    if (tid < BW)
      prow[tid] = tid == k ? rp : Pn[(size_t)p * ldp + tid] * rp;
    __syncthreads();
    // ---- eliminate column k from every other row; the column itself receives the new column of the inverse
    // (row[k] <- 0 first, so that the uniform update row[c] -= f prow[c] leaves -f / pivot there); two values per access
    for (int i = tid; i < n; i += TPB)
    {
      V* row = Pn + (size_t)i * ldp;
      if (i == p)
      {
#pragma unroll
        for (int c = 0; c < BW; c += 2)
          st2(row + c, prow[c], prow[c + 1]);
      }
      else
      {
        const V f = row[k];
        row[k]    = V(0.0);
#pragma unroll
        for (int c = 0; c < BW; c += 2)
        {
          V a0, a1, p0, p1;
          ld2(row + c, a0, a1);
          ld2(prow + c, p0, p1);
          st2(row + c, a0 - f * p0, a1 - f * p1);
        }
      }
    }
    __syncthreads();
  }

  // ---- results: panel columns of X <- E + W; W (zero beyond b); pivot rows of the OLD matrix outside the panel
  for (int e = tid; e < n * BW; e += TPB)
  {
    const int i = e / BW, k = e - i * BW;
    V w = V(0.0);
    if (k < b)
    {
      const V v                     = Pn[(size_t)i * ldp + k];
      Xw[(size_t)i * n + j0 + k]    = v;
      w                             = i == Rs[k] ? v - V(1.0) : v;
    }
    W[((size_t)iw * n + i) * BW + k] = w;
  }
  for (int e = tid; e < BW * n; e += TPB)
  {
    const int k = e / n, c = e - k * n;
    V v = V(0.0);
    if (k < b && (c < j0 || c >= j0 + b))
      v = Xw[(size_t)Rs[k] * n + c];
    Xr[((size_t)iw * BW + k) * n + c] = v;
  }
  for (int i = tid; i < n; i += TPB)
    ug[i] = us[i];
  if (tid < b)
    R[(size_t)iw * n + j0 + tid] = Rs[tid];
  if (tid == 0)
  {
    if (j0 == 0)
    {
      logdet[2 * (size_t)iw]     = ld[0];
      logdet[2 * (size_t)iw + 1] = ld[1];
      info[iw]                   = bad;
    }
    else
    {
      logdet[2 * (size_t)iw] += ld[0];
      logdet[2 * (size_t)iw + 1] += ld[1];
      if (bad && info[iw] == 0)
        info[iw] = bad;
    }
  }
}

// X[:, c] += W X[R, c] for every column outside the panel [j0, j0 + b).  grid (row tiles, walkers); one CTA owns a 64-row
// block of one walker's matrix and walks over its 64-column tiles: the W block (64 x BW) is staged once, the X[R, :]
// tiles (BW x 64) stream through a double-buffered cp.async ring, and the accumulator tile of the NEXT column tile is
// loaded into registers while the DMMAs of the current one run -- the kernel is a stream over X (one read, one write per
// panel step), so what matters is that every CTA always has loads in flight.  Warp (wr, wc) owns 16 rows x 32 columns
// = 2 x 4 accumulator tiles.
template<typename V, int BW>
__global__ void __launch_bounds__(TPB) gj_update_kernel(V* __restrict__ X, const int n, const int j0, const int b,
                                                        const V* __restrict__ W, const V* __restrict__ Xr)
{
  using namespace wb64;
  constexpr int SW = BW + 4, SX = TILE + 4; // conflict-free fragment loads (woodbury_dmma.cuh)
  extern __shared__ __align__(16) unsigned char gj_smem[];
  V* Ws = reinterpret_cast<V*>(gj_smem);
  V* Xs = Ws + TILE * SW; // [2][BW][SX]
  const int iw = blockIdx.y, r0 = blockIdx.x * TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int nt = (n + TILE - 1) / TILE;
  V* Xw        = X + (size_t)iw * n * n;
  const V* Xrw = Xr + (size_t)iw * BW * n;
  const int wr = warp >> 1, wc = warp & 1;
  // a column tile that lies entirely inside the panel has nothing to do
  auto skip = [&](int ct) { return ct * TILE >= j0 && ct * TILE + TILE <= j0 + b; };
  auto next_tile = [&](int ct) {
    ++ct;
    while (ct < nt && skip(ct))
      ++ct;
    return ct;
  };
  // X[R, tile] -> shared memory (asynchronous; columns beyond n are zero)
  auto stage_x = [&](int ct, int buf) {
    V* dst = Xs + (size_t)buf * BW * SX;
    for (int e = tid; e < BW * TILE; e += TPB)
    {
      const int k = e / TILE, c = e - k * TILE, col = ct * TILE + c;
      if (col < n)
      {
        if constexpr (sizeof(V) == 8)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + k * SX + c)),
                       "l"(Xrw + (size_t)k * n + col)
                       : "memory");
        else
          cp_async16(dst + k * SX + c, Xrw + (size_t)k * n + col);
      }
      else
        dst[k * SX + c] = zero_v<V>();
    }
    cp_async_commit();
  };
  // (real values: the next tile's accumulators wait in registers; complex values need twice the registers per tile
  // and load theirs at the top of the tile's own iteration instead)
  constexpr bool AHEAD = !value_traits<V>::is_complex;
  Acc<V> acc[2][4], nxt[AHEAD ? 2 : 1][AHEAD ? 4 : 1];
  auto load_c = [&](int ct, auto& dst) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        const int row = r0 + 16 * wr + 8 * i + g, col = ct * TILE + 32 * wc + 8 * j + 2 * t;
        dst[i][j].set(0, (row < n && col < n) ? Xw[(size_t)row * n + col] : zero_v<V>());
        dst[i][j].set(1, (row < n && col + 1 < n) ? Xw[(size_t)row * n + col + 1] : zero_v<V>());
      }
  };

  int ct = next_tile(-1);
  if (ct >= nt)
    return;
  stage_x(ct, 0);
  if constexpr (AHEAD)
    load_c(ct, nxt);
  for (int e = tid; e < TILE * BW; e += TPB)
  {
    const int i = e / BW, k = e - i * BW, row = r0 + i;
    Ws[i * SW + k] = row < n ? W[((size_t)iw * n + row) * BW + k] : zero_v<V>();
  }
  int buf = 0;
  while (ct < nt)
  {
    cp_async_wait<0>();
    __syncthreads(); // this tile's X[R, :] (and, the first time, W) is in shared memory; the other buffer is free
    if constexpr (AHEAD)
    {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = nxt[i][j];
    }
    else
      load_c(ct, acc);
    const int cn = next_tile(ct);
    if (cn < nt)
    {
      stage_x(cn, buf ^ 1);
      if constexpr (AHEAD)
        load_c(cn, nxt);
    }
    const V* Xb = Xs + (size_t)buf * BW * SX;
#pragma unroll
    for (int ks = 0; ks < BW / 4; ++ks)
    {
      V a[2], bq[4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        a[i] = lds_v(Ws + (16 * wr + 8 * i + g) * SW + 4 * ks + t);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        bq[j] = lds_v(Xb + (4 * ks + t) * SX + 32 * wc + 8 * j + g);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j].mma(a[i], bq[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        const int row = r0 + 16 * wr + 8 * i + g, col = ct * TILE + 32 * wc + 8 * j + 2 * t;
        if (row < n)
        {
          if (col < n && (col < j0 || col >= j0 + b))
            Xw[(size_t)row * n + col] = acc[i][j].get(0);
          if (col + 1 < n && (col + 1 < j0 || col + 1 >= j0 + b))
            Xw[(size_t)row * n + col + 1] = acc[i][j].get(1);
        }
      }
    ct = cn;
    buf ^= 1;
  }
}

// inverse of the pivot permutation and its parity (one CTA per walker): Rinv[R[c]] = c; the sign of the permutation
// enters the phase of the log-determinant
__global__ void gj_perm_kernel(const int* __restrict__ R, int* __restrict__ Rinv, const int n, double* __restrict__ logdet)
{
  extern __shared__ __align__(16) unsigned char gj_smem[];
  unsigned char* seen = gj_smem;
  const int iw = blockIdx.x, tid = threadIdx.x;
  const int* Rw = R + (size_t)iw * n;
  for (int c = tid; c < n; c += blockDim.x)
  {
    Rinv[(size_t)iw * n + Rw[c]] = c;
    seen[c]                      = 0;
  }
  __syncthreads();
  if (tid == 0)
  {
    int transpositions = 0;
    for (int i = 0; i < n; ++i)
      if (!seen[i])
      {
        int len = 0;
        for (int j = i; !seen[j]; j = Rw[j])
          seen[j] = 1, ++len;
        transpositions += len - 1;
      }
    if (transpositions & 1)
      logdet[2 * (size_t)iw + 1] += 3.14159265358979323846;
  }
}

// Ainv[c][j] = S[R[c]][Rinv[j]] cast to the determinant's value type
template<typename T>
__global__ void gj_store_inverse_kernel(const DetDev<T> D, const typename value_traits<T>::dbl_t* __restrict__ S,
                                        const int* __restrict__ R, const int* __restrict__ Rinv)
{
  const int iw = blockIdx.z, c = blockIdx.y, n = D.n;
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n)
    D.Ainv[((size_t)iw * n + c) * D.lda + j] =
        from_dbl<T>(S[((size_t)iw * n + R[(size_t)iw * n + c]) * n + Rinv[(size_t)iw * n + j]]);
}
} // namespace gj
#endif
} // namespace qmcb
