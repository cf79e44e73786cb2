// qmcpack_b200/csrc/woodbury.cuh -- rank-k Woodbury flush of the inverse Slater matrices on tensor cores (sm_100a).
//
// Replaces DelayedUpdateBatched::mw_updateInvMat (QMCWaveFunctions/Fermion/DelayedUpdateBatched.h:675-738): three
// cuBLAS gemmBatched calls + applyW_batched, i.e. three passes over Ainv.  Math (row-major):
//     T[m][a]  = sum_j Ainv[m][j] U[a][j]  - (m == list[a])        (n x c)
//     U'[a][j] = sum_b Binv[a][b] V[b][j]                           (c x n)
//     Ainv[m][j] -= sum_a T[m][a] U'[a][j]
// Here ONE pass: a CTA owns row tiles of one walker's Ainv.  U and U' stay in shared memory for the whole CTA; each
// 32-row tile of Ainv is staged once with cp.async (double buffered: the next tile streams in while the current one is
// on the tensor cores), T = tile * U^T is formed on the tensor cores, fixed up, and the tile is updated by a second
// tensor-core product and written straight back to HBM from the accumulator fragments (full 32-byte sectors).
// Ainv traffic = one read + one write (the reference: 3 reads + 1 write + the tempMat round trip).
//
// float path: mma.sync.m16n8k8 TF32 with the 3-product split (x = hi + lo, D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi) so the
// update keeps FP32-level accuracy -- plain TF32 (10-bit mantissa) would destroy the inverse.  The three products go to
// independent accumulators (summed at the end) so that the dependent-MMA chain is one third as long.
#pragma once
#include "common.cuh"
#include "det.cuh"

namespace qmcb
{
#ifdef __CUDACC__
namespace wb
{
constexpr int RT  = 32;  // Ainv rows per tile
constexpr int KD  = 32;  // delay slots handled (c <= KD; unused slots are zero rows)
constexpr int TPB = 512; // 16 warps: one CTA per SM (208 KB of shared memory), so the warps must come from the CTA

// x = hi + lo with hi the TF32 truncation of x (mask, one LOP3) and lo = x - hi (exact in FP32; the tensor core reads
// its top 19 bits).  cvt.rna.tf32 would round instead of truncate but runs on the slow conversion pipe: with two cvt per
// operand element the split, not the MMA, bounded the kernel (measured: 803 us -> see DESIGN.md).  Truncation leaves a
// relative error of 2^-20 per product, 16x below what the float inverse itself carries.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// shared-memory row strides (in floats) chosen so that the fragment loads are bank-conflict free:
//   A-type access  M[g][t]   (8 rows x 4 cols): stride = 4 (mod 8)
//   B-type access  M[t][g]   (4 rows x 8 cols): stride = 8 (mod 16)
__host__ __device__ inline int stride_a(int n) { return ((n + 7) / 8) * 8 + 4; }
__host__ __device__ inline int stride_b(int n) { return ((n + 7) / 8) * 8 + 8; }

inline size_t smem_bytes_f32(int n)
{
  return sizeof(float) * ((size_t)KD * stride_a(n)       /* Us   [KD][sa]     */
                          + (size_t)KD * stride_b(n)     /* Ups  [KD][sb]     */
                          + (size_t)2 * RT * stride_a(n) /* At   [2][RT][sa]  */
                          + (size_t)2 * RT * (KD + 4))   /* Ts   [2][RT][KD+4] (one per K half) */
      + KD * sizeof(int);
}

// stage `rows` rows of `n` floats (global row stride ld) into smem rows of stride `ss`; rows >= nvalid and the pad
// columns [n, n8) are zero-filled with plain stores, the body goes through cp.async (n % 4 == 0 and 16-byte aligned rows)
__device__ __forceinline__ void stage_rows_async(float* dst, int ss, const float* src, int ld, int rows, int nvalid, int n,
                                                 int n8, int tid)
{
  const int n4 = n8 / 4;
  for (int e = tid; e < rows * n4; e += TPB)
  {
    const int r = e / n4, j4 = (e - r * n4) * 4;
    float* d = dst + (size_t)r * ss + j4;
    if (r < nvalid && j4 + 3 < n)
      cp_async16(d, src + (size_t)r * ld + j4);
    else
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid)
      {
        const float* s = src + (size_t)r * ld + j4;
        if (j4 < n) v.x = s[0];
        if (j4 + 1 < n) v.y = s[1];
        if (j4 + 2 < n) v.z = s[2];
      }
      *reinterpret_cast<float4*>(d) = v;
    }
  }
}

// grid = (nw, S): CTA (iw, s) updates row tiles s, s+S, ... of walker iw.  c <= KD pending delays.
// requires n % 4 == 0 (rows of Ainv, U, V 16-byte aligned for cp.async); the host falls back otherwise
__global__ void __launch_bounds__(TPB, 1) woodbury_flush_tf32_kernel(const DetDev<float> D, const int c)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = D.n, lda = D.lda, k = D.k;
  const int n8 = ((n + 7) / 8) * 8;
  const int sa = stride_a(n), sb = stride_b(n), st = KD + 4;
  float* Us  = reinterpret_cast<float*>(smem_raw);
  float* Ups = Us + (size_t)KD * sa;
  float* At0 = Ups + (size_t)KD * sb;
  float* Ts  = At0 + (size_t)2 * RT * sa;
  int* lst   = reinterpret_cast<int*>(Ts + (size_t)2 * RT * st);
  const int iw = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* U = D.U + (size_t)iw * k * n;
  const float* V = D.V + (size_t)iw * k * n;
  const float* B = D.Binv + (size_t)iw * k * k;
  float* Ainv    = D.Ainv + (size_t)iw * n * lda;
  const int ntiles = (n + RT - 1) / RT;

  // ---- stage U -> Us, V -> At buffer 1 (scratch; KD == RT rows), Binv -> Ts (scratch), list; unused slots are zero
  stage_rows_async(Us, sa, U, n, KD, c, n, n8, tid);
  stage_rows_async(At0 + (size_t)RT * sa, sa, V, n, KD, c, n, n8, tid);
  cp_async_commit();
  // first Ainv tile of this CTA streams in behind them
  int tile = blockIdx.y;
  if (tile < ntiles)
    stage_rows_async(At0, sa, Ainv + (size_t)tile * RT * lda, lda, RT, n - tile * RT, n, n8, tid);
  cp_async_commit();
  for (int e = tid; e < KD * KD; e += TPB)
  {
    const int a = e / KD, b = e - a * KD;
    Ts[a * st + b] = (a < c && b < c) ? B[a * k + b] : 0.f;
  }
  if (tid < KD)
    lst[tid] = tid < c ? D.list[(size_t)iw * k + tid] : -1;
  cp_async_wait<1>(); // U and V landed (the Ainv tile may still be in flight)
  __syncthreads();
  // ---- U' = Binv * V  (KD x n8, K = KD): register-tiled FMAs (8 slots per thread), 0.8 MFLOP
  {
    const float* Vs = At0 + (size_t)RT * sa;
    for (int e = tid; e < (KD / 8) * n8; e += TPB)
    {
      const int ag = e / n8, j = e - ag * n8;
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int b = 0; b < KD; ++b)
      {
        const float v = Vs[(size_t)b * sa + j];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          s[q] += Ts[(ag * 8 + q) * st + b] * v;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        Ups[(size_t)(ag * 8 + q) * sb + j] = s[q];
    }
  }
  __syncthreads();

  int buf = 0;
  for (; tile < ntiles; tile += gridDim.y, buf ^= 1)
  {
    const int m0 = tile * RT;
    float* At    = At0 + (size_t)buf * RT * sa;
    // prefetch the next tile into the other buffer, then wait for the current one
    const int next = tile + gridDim.y;
    if (next < ntiles)
      stage_rows_async(At0 + (size_t)(buf ^ 1) * RT * sa, sa, Ainv + (size_t)next * RT * lda, lda, RT, n - next * RT, n, n8,
                       tid);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // ---- GEMM 1: T[RT x KD] = At[RT x n8] * Us^T ; 2 (rows) x 4 (slots) x 2 (K halves) warps, one 16 x 8 tile each,
    //      three independent accumulator chains (one per split product); the K halves are added when GEMM 2 loads T
    {
      const int r0 = (warp & 1) * 16, a0 = ((warp >> 1) & 3) * 8, kh = warp >> 3;
      const int kmid = ((n8 / 8 + 1) / 2) * 8;
      const int kbeg = kh ? kmid : 0, kend = kh ? n8 : kmid;
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
      const float* Ar0 = At + (size_t)(r0 + g) * sa + t;
      const float* Ar1 = Ar0 + 8 * sa;
      const float* Ub  = Us + (size_t)(a0 + g) * sa + t;
#pragma unroll 4
      for (int kk = kbeg; kk < kend; kk += 8)
      {
        uint32_t ah[4], al[4], bh[2], bl[2];
        split_tf32(Ar0[kk], ah[0], al[0]);
        split_tf32(Ar1[kk], ah[1], al[1]);
        split_tf32(Ar0[kk + 4], ah[2], al[2]);
        split_tf32(Ar1[kk + 4], ah[3], al[3]);
        split_tf32(Ub[kk], bh[0], bl[0]);
        split_tf32(Ub[kk + 4], bh[1], bl[1]);
        mma_tf32(acc0, al, bh);
        mma_tf32(acc1, ah, bl);
        mma_tf32(acc2, ah, bh);
      }
      // store -T (the second product then ADDS (-T) * U'); small terms first
      float* d0 = Ts + (size_t)kh * RT * st + (size_t)(r0 + g) * st + a0 + 2 * t;
      d0[0]          = -((acc0[0] + acc1[0]) + acc2[0]);
      d0[1]          = -((acc0[1] + acc1[1]) + acc2[1]);
      d0[8 * st]     = -((acc0[2] + acc1[2]) + acc2[2]);
      d0[8 * st + 1] = -((acc0[3] + acc1[3]) + acc2[3]);
    }
    __syncthreads();
    // applyW: T[list[a]][a] -= 1  -> (-T) += 1 ; pseudo-accepted slots carry -1 and are skipped
    if (tid < KD)
    {
      const int m = lst[tid] - m0;
      if (lst[tid] >= 0 && m >= 0 && m < RT)
        Ts[(size_t)m * st + tid] += 1.f;
    }
    __syncthreads();

    // ---- GEMM 2: tile += (-T)[RT x KD] * U'[KD x n8] ; every warp owns all 32 rows x (n8/8)/16 column tiles
    {
      const int ntw = (n8 / 8 + 15) / 16; // 8-column tiles per warp (3 for n = 384)
      const int j0  = warp * ntw * 8;
      uint32_t ah[2][4][4], al[2][4][4]; // [m-tile][k-step][frag]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
        {
          const float* T0 = Ts + (size_t)(mt * 16 + g) * st + ks * 8 + t;
          const float* T1 = T0 + (size_t)RT * st; // second K half
          split_tf32(T0[0] + T1[0], ah[mt][ks][0], al[mt][ks][0]);
          split_tf32(T0[8 * st] + T1[8 * st], ah[mt][ks][1], al[mt][ks][1]);
          split_tf32(T0[4] + T1[4], ah[mt][ks][2], al[mt][ks][2]);
          split_tf32(T0[8 * st + 4] + T1[8 * st + 4], ah[mt][ks][3], al[mt][ks][3]);
        }
      for (int nt = 0; nt < ntw; ++nt)
      {
        const int jc = j0 + nt * 8;
        if (jc >= n8)
          break;
        float acc[2][4], lo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
          const float* C0 = At + (size_t)(mt * 16 + g) * sa + jc + 2 * t;
          acc[mt][0]      = C0[0];
          acc[mt][1]      = C0[1];
          acc[mt][2]      = C0[8 * sa];
          acc[mt][3]      = C0[8 * sa + 1];
          lo[mt][0] = lo[mt][1] = lo[mt][2] = lo[mt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
        {
          uint32_t bh[2], bl[2];
          const float* P0 = Ups + (size_t)(ks * 8 + t) * sb + jc + g;
          split_tf32(P0[0], bh[0], bl[0]);
          split_tf32(P0[4 * sb], bh[1], bl[1]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
          {
            mma_tf32(lo[mt], al[mt][ks], bh);
            mma_tf32(lo[mt], ah[mt][ks], bl);
            mma_tf32(acc[mt], ah[mt][ks], bh);
          }
        }
        // straight to HBM from the fragments: every (row, 8 columns) is one full 32-byte sector
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
          const int row = m0 + mt * 16 + g, col = jc + 2 * t;
          if (row < n)
          {
            float* dst = Ainv + (size_t)row * lda + col;
            if (col + 1 < n)
              *reinterpret_cast<float2*>(dst) = make_float2(acc[mt][0] + lo[mt][0], acc[mt][1] + lo[mt][1]);
            else if (col < n)
              dst[0] = acc[mt][0] + lo[mt][0];
          }
          if (row + 8 < n)
          {
            float* dst = Ainv + (size_t)(row + 8) * lda + col;
            if (col + 1 < n)
              *reinterpret_cast<float2*>(dst) = make_float2(acc[mt][2] + lo[mt][2], acc[mt][3] + lo[mt][3]);
            else if (col < n)
              dst[0] = acc[mt][2] + lo[mt][2];
          }
        }
      }
    }
    __syncthreads(); // this buffer and Ts are overwritten by the next iterations
  }
  cp_async_wait<0>();
}
} // namespace wb
#endif
} // namespace qmcb
