// qmcpack_b200/csrc/woodbury_dmma.cuh -- rank-k Woodbury flush in FULL precision on the FP64 tensor pipe (sm_100a).
//
// Same contraction as woodbury.cuh / woodbury_tc5.cuh (DelayedUpdateBatched::mw_updateInvMat,
// Fermion/DelayedUpdateBatched.h:675-738; CPU form Fermion/DelayedUpdate.h:145-217), for double and complex<double>
// inverses (the reference's full-precision and QMC_COMPLEX builds, the NiO-a128 class):
//     T[m][a]  = sum_j Ainv[m][j] U[a][j]  - (m == list[a])        (n x c)
//     U'[a][j] = sum_b Binv[a][b] V[b][j]                           (c x n)   <- small SIMT GEMM, launched before
//     Ainv[m][j] -= sum_a T[m][a] U'[a][j]
// The reference streams Ainv three times through cuBLAS gemmBatched.  Here ONE pass with mma.sync.m8n8k4.f64 (SASS DMMA):
//   * a CTA (8 warps) owns 64-row tiles of one walker's Ainv.  The tile itself is never staged in shared memory: the A
//     fragments of the first product (8 rows x 4 consecutive columns per warp load = full 32-byte sectors) go from global
//     memory straight into registers one K chunk ahead, and in the second product the tile is the ACCUMULATOR, loaded
//     from L2 one chunk ahead and stored back from the fragments (64-byte row segments).
//   * U and U' stream through a cp.async ring of K chunks (STAGES deep) as one continuous sequence over both products and
//     all tiles of the CTA, so the pipeline never drains between the products; one __syncthreads per chunk.
//   * warp w owns rows 8w..8w+7 of the tile in the first product (all KD slots: one A fragment feeds KD/8 DMMAs) and
//     writes -T, with the `-1` fix-up of applyW_batched, to shared memory; in the second product a warp owns a
//     16 x (KC/2) block (2 row tiles x KC/16 column tiles).
//   * complex: every thread holds the (re, im) pair of its fragment element, so one complex product is four DMMAs on the
//     same fragment layout (re*re, -im*im -> real accumulator; re*im, im*re -> imaginary accumulator); plain (unconjugated)
//     products exactly as the reference's gemm('T','N').
// Shared-memory row strides are chosen so that every fragment load is bank-conflict free (64-bit accesses are served per
// half warp, 128-bit accesses per quarter warp).
#pragma once
#include "common.cuh"
#include "det.cuh"

namespace qmcb
{
#ifdef __CUDACC__
namespace wb64
{
constexpr int RT  = 64;  // Ainv rows per tile (8 warps x 8 rows)
constexpr int TPB = 256;

__device__ __forceinline__ void dmma(double (&d)[2], const double a, const double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
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

// L2 eviction policies: the tile is read twice (A operand of the first product, accumulator of the second) about twelve
// chunk steps apart with ~50 MB of other traffic in between -- ncu showed the second read going back to DRAM.  The first
// read marks the lines evict_last, the second read and the write-back mark them evict_first.
__device__ __forceinline__ uint64_t policy_evict_last()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_hint(const double* p, const uint64_t pol)
{
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ cx<double> ld_hint(const cx<double>* p, const uint64_t pol)
{
  double x, y;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(x), "=d"(y) : "l"(p), "l"(pol));
  return cx<double>(x, y);
}
__device__ __forceinline__ double2 ld_hint2(const double* p, const uint64_t pol)
{
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint2(double* p, const double x, const double y, const uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(x), "d"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk_hint(const void* p, unsigned bytes, const uint64_t pol)
{
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(pol) : "memory");
}

// one 8 x 8 accumulator tile of the m8n8k4 shape: the thread (g = lane / 4, t = lane % 4) holds (g, 2t) and (g, 2t + 1)
template<typename V>
struct Acc;
template<>
struct Acc<double>
{
  double d[2];
  __device__ __forceinline__ void zero() { d[0] = d[1] = 0.0; }
  __device__ __forceinline__ void mma(const double a, const double b) { dmma(d, a, b); }
  __device__ __forceinline__ double get(int i) const { return d[i]; }
  __device__ __forceinline__ void set(int i, const double v) { d[i] = v; }
};
template<>
struct Acc<cx<double>>
{
  double re[2], im[2];
  __device__ __forceinline__ void zero() { re[0] = re[1] = im[0] = im[1] = 0.0; }
  __device__ __forceinline__ void mma(const cx<double>& a, const cx<double>& b)
  {
    dmma(re, a.re, b.re);
    dmma(re, -a.im, b.im);
    dmma(im, a.re, b.im);
    dmma(im, a.im, b.re);
  }
  __device__ __forceinline__ cx<double> get(int i) const { return cx<double>(re[i], im[i]); }
  __device__ __forceinline__ void set(int i, const cx<double>& v)
  {
    re[i] = v.re;
    im[i] = v.im;
  }
};

template<typename V>
__device__ __forceinline__ V zero_v();
template<>
__device__ __forceinline__ double zero_v<double>()
{
  return 0.0;
}
template<>
__device__ __forceinline__ cx<double> zero_v<cx<double>>()
{
  return cx<double>(0.0, 0.0);
}
__device__ __forceinline__ double lds_v(const double* p) { return *p; }
__device__ __forceinline__ cx<double> lds_v(const cx<double>* p)
{
  const double2 v = *reinterpret_cast<const double2*>(p);
  return cx<double>(v.x, v.y);
}

template<typename V>
struct Cfg
{
  static constexpr bool CPLX = value_traits<V>::is_complex;
  static constexpr int EPP   = CPLX ? 1 : 2; // elements per 16-byte cp.async piece
  // row strides (in elements) of the shared-memory chunks; see the header comment
  __host__ __device__ static constexpr int su(int KC) { return KC + 4; }                // A-type access  M[g][t]
  __host__ __device__ static constexpr int sp(int KC) { return CPLX ? KC + 2 : KC + 4; } // B-type access  M[t][g]
  __host__ __device__ static constexpr int st(int KD) { return KD + 4; }                // -T, A-type access
};

template<typename V, int KD, int KC, int STAGES>
constexpr size_t smem_bytes()
{
  return sizeof(V) * ((size_t)STAGES * KD * Cfg<V>::su(KC) + (size_t)RT * Cfg<V>::st(KD)) + KD * sizeof(int);
}

// U'[a][j] = sum_b Binv[a][b] V[b][j]  (c x n, K = c): one thread per column j and block of eight slots a
// (grid = (ceil(n / 128), nw, ceil(c / 8)), 128 threads); the eight Binv rows sit transposed in shared memory so that the
// factors of one b are broadcast 16-byte loads; V is read coalesced, eight loads in flight per thread.
// 0.8 MFLOP per walker at a64: this kernel only has to stay out of the way.
template<typename V>
__global__ void __launch_bounds__(128) binv_v_kernel(const DetDev<V> D, const int c)
{
  __shared__ __align__(16) V Bt[64 * 8]; // Bt[b][q] = Binv[a0 + q][b]
  const int n = D.n, k = D.k, iw = blockIdx.y, a0 = blockIdx.z * 8;
  const V* B = D.Binv + (size_t)iw * k * k;
  for (int e = threadIdx.x; e < 8 * c; e += blockDim.x)
  {
    const int q = e / c, b = e - q * c;
    Bt[b * 8 + q] = a0 + q < c ? B[(size_t)(a0 + q) * k + b] : zero_v<V>();
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  const V* Vw = D.V + (size_t)iw * k * n + j;
  V* Uw       = D.Up + (size_t)iw * k * n + j;
  V acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    acc[q] = zero_v<V>();
#pragma unroll 8
  for (int b = 0; b < c; ++b)
  {
    const V v = Vw[(size_t)b * n];
#pragma unroll
    for (int q = 0; q < 8; ++q)
      acc[q] += Bt[b * 8 + q] * v;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (a0 + q < c)
      Uw[(size_t)(a0 + q) * n] = acc[q];
}

// The same product on the FP64 tensor pipe: one CTA (8 warps) per walker, Binv staged once in shared memory (A operand,
// row stride = 4 mod 16 doubles), the B fragments of a column tile (4 rows x 8 consecutive columns of V = 64-byte
// segments) straight from global memory one tile ahead, U' stored from the accumulator fragments.  No shared memory for V.
// grid = nw, 256 threads, KD * (KD + 4) * sizeof(V) bytes of dynamic shared memory.
template<typename V, int KD>
__global__ void __launch_bounds__(TPB) binv_v_dmma_kernel(const DetDev<V> D, const int c)
{
  constexpr int SB = KD + 4, NR = KD / 8, KS = KD / 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* Bs = reinterpret_cast<V*>(smem_raw);
  const int n = D.n, k = D.k, iw = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const V* B  = D.Binv + (size_t)iw * k * k;
  const V* Vw = D.V + (size_t)iw * k * n;
  V* Uw       = D.Up + (size_t)iw * k * n;
  for (int e = tid; e < KD * KD; e += TPB)
  {
    const int a = e / KD, b = e - a * KD;
    Bs[a * SB + b] = (a < c && b < c) ? B[(size_t)a * k + b] : zero_v<V>();
  }
  const int ntile = (n + 7) / 8;
  auto load_b = [&](const int ct, V (&b)[KS]) {
    const int col = ct * 8 + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
    {
      const int row = ks * 4 + t;
      b[ks]         = (ct < ntile && row < c && col < n) ? Vw[(size_t)row * n + col] : zero_v<V>();
    }
  };
  V bcur[KS], bnext[KS];
  load_b(warp, bnext);
  __syncthreads();
  for (int ct = warp; ct < ntile; ct += TPB / 32)
  {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      bcur[ks] = bnext[ks];
    load_b(ct + TPB / 32, bnext);
#pragma unroll
    for (int rt = 0; rt < NR; ++rt)
    {
      if (rt * 8 >= c)
        break;
      Acc<V> acc;
      acc.zero();
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        acc.mma(lds_v(Bs + (rt * 8 + g) * SB + ks * 4 + t), bcur[ks]);
      const int row = rt * 8 + g, col = ct * 8 + 2 * t;
      if (row < c)
      {
        if (col < n)
          Uw[(size_t)row * n + col] = acc.get(0);
        if (col + 1 < n)
          Uw[(size_t)row * n + col + 1] = acc.get(1);
      }
    }
  }
}

// real: rows of U, U' (stride n) must be 16-byte aligned for cp.async, i.e. n even
template<typename V>
inline bool eligible(int n, int k, int c, int KD)
{
  return k <= KD && c <= KD && n >= 8 && (value_traits<V>::is_complex || n % 2 == 0);
}

// grid = (S, nw): CTA (s, iw) updates row tiles s, s+S, ... of walker iw (tile index fastest, so that the CTAs resident at
// any time belong to few walkers and share their U, U' chunks in L2).  c <= KD pending delays; D.Up holds U' = Binv*V.
template<typename V, int KD, int KC, int STAGES>
__global__ void __launch_bounds__(TPB, value_traits<V>::is_complex ? 1 : 2)
    woodbury_flush_dmma_kernel(const DetDev<V> D, const int c)
{
  using C = Cfg<V>;
  static_assert(KC % 16 == 0 && KD % 8 == 0, "chunk shapes");
  constexpr int SU = C::su(KC), SP = C::sp(KC), ST = C::st(KD);
  constexpr int STAGE_ELEMS = KD * SU; // SP <= SU
  constexpr int NT1 = KD / 8;          // slot tiles of the first product
  constexpr int KS1 = KC / 4;          // k-steps per chunk, first product
  constexpr int KS2 = KD / 4;          // k-steps, second product
  constexpr int JT  = KC / 16;         // column tiles per warp and chunk, second product
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* ring  = reinterpret_cast<V*>(smem_raw);
  V* Ts    = ring + (size_t)STAGES * STAGE_ELEMS;
  int* lst = reinterpret_cast<int*>(Ts + (size_t)RT * ST);

  const int n = D.n, lda = D.lda, k = D.k;
  const int iw = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const V* U  = D.U + (size_t)iw * k * n;
  const V* Up = D.Up + (size_t)iw * k * n;
  V* Ainv     = D.Ainv + (size_t)iw * n * lda;
  const int ntiles = (n + RT - 1) / RT;
  const int nch    = (n + KC - 1) / KC;
  const int niter  = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total  = niter * 2 * nch;
  if (total == 0)
    return;
  const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
  if (tid == 0)
  {
    const int m0 = blockIdx.x * RT;
    prefetch_l2_bulk_hint(Ainv + (size_t)m0 * lda, (unsigned)(min(RT, n - m0) * lda * sizeof(V)), pol_keep);
  }
  if (tid < KD)
    lst[tid] = tid < c ? D.list[(size_t)iw * k + tid] : -1;

  // chunk s of the stream: first product chunks (U columns) then second product chunks (U' columns) of every tile
  auto stage = [&](const int s) {
    const int q     = s % (2 * nch);
    const bool sec  = q >= nch;
    const int col0  = (sec ? q - nch : q) * KC;
    const V* src    = sec ? Up : U;
    const int ss    = sec ? SP : SU;
    V* dst          = ring + (size_t)(s % STAGES) * STAGE_ELEMS;
    constexpr int P = KC / C::EPP; // pieces per row
    for (int e = tid; e < KD * P; e += TPB)
    {
      const int a = e / P, j = (e - a * P) * C::EPP;
      V* d = dst + a * ss + j;
      if (a < c && col0 + j < n)
        cp_async16(d, src + (size_t)a * n + col0 + j);
      else
        *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
    }
  };

  // A fragments of the first product for chunk q of the tile starting at row m0: element (m0 + 8*warp + g, q*KC + 4*ks + t)
  V anext[KS1];
  auto load_a = [&](const int m0, const int q) {
    const int row = m0 + warp * 8 + g;
    const V* p    = Ainv + (size_t)row * lda + q * KC + t;
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks)
      anext[ks] = (row < n && q * KC + ks * 4 + t < n) ? ld_hint(p + ks * 4, pol_keep) : zero_v<V>();
  };
  // accumulator fragments of the second product: warp block = row tiles 2*rp + {0,1}, column tiles cp*JT + {0..JT-1}
  const int rp = warp & 3, cp = warp >> 2;
  V cnext[2][JT][2];
  auto load_c = [&](const int m0, const int j0) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < JT; ++j)
      {
        const int row = m0 + (2 * rp + i) * 8 + g, col = j0 + (cp * JT + j) * 8 + 2 * t;
        const V* p    = Ainv + (size_t)row * lda + col;
        if (row < n && col + 1 < n)
        {
          if constexpr (C::CPLX)
          {
            cnext[i][j][0] = ld_hint(p, pol_drop);
            cnext[i][j][1] = ld_hint(p + 1, pol_drop);
          }
          else
          {
            const double2 v = ld_hint2(p, pol_drop);
            cnext[i][j][0]  = v.x;
            cnext[i][j][1]  = v.y;
          }
        }
        else
        {
          cnext[i][j][0] = (row < n && col < n) ? lds_v(p) : zero_v<V>();
          cnext[i][j][1] = zero_v<V>();
        }
      }
  };

  for (int s = 0; s < STAGES - 1; ++s)
  {
    if (s < total)
      stage(s);
    cp_async_commit();
  }
  load_a(blockIdx.x * RT, 0);

  // one pipeline step: chunk s has landed and is visible, the buffer consumed at step s - 1 is refilled
  auto step = [&](const int s) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (s + STAGES - 1 < total)
      stage(s + STAGES - 1);
    cp_async_commit();
  };
  // real, k <= 32: the -T fragments of the second product (2 row tiles x 8 k-steps = 16 doubles) stay in registers for
  // the whole tile; otherwise they are re-read from shared memory per chunk (complex: four DMMAs per fragment pair)
  constexpr bool HOIST = !C::CPLX && KD <= 32;
  int s = 0;
  for (int it = 0; it < niter; ++it)
  {
    const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * RT;
    {
      const int m0n = m0 + (int)gridDim.x * RT;
      if (tid == 0 && m0n < n)
        prefetch_l2_bulk_hint(Ainv + (size_t)m0n * lda, (unsigned)(min(RT, n - m0n) * lda * sizeof(V)), pol_keep);
    }
    // ---- first product: -T[8 rows of this warp][KD] accumulates over the K chunks
    {
      Acc<V> acc1[NT1];
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt)
        acc1[nt].zero();
      for (int q = 0; q < nch; ++q, ++s)
      {
        step(s);
        const V* buf = ring + (size_t)(s % STAGES) * STAGE_ELEMS;
        V acur[KS1];
#pragma unroll
        for (int ks = 0; ks < KS1; ++ks)
          acur[ks] = anext[ks];
        if (q + 1 < nch)
          load_a(m0, q + 1);
        else
          load_c(m0, 0);
#pragma unroll
        for (int ks = 0; ks < KS1; ++ks)
#pragma unroll
          for (int nt = 0; nt < NT1; ++nt)
            acc1[nt].mma(acur[ks], lds_v(buf + (nt * 8 + g) * SU + ks * 4 + t));
      }
      // store -T with the applyW fix-up ( T[list[a]][a] -= 1  ->  (-T) += 1 ); pseudo-accepted slots carry -1
      const int row = m0 + warp * 8 + g;
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i)
        {
          const int a = nt * 8 + 2 * t + i;
          V v         = -acc1[nt].get(i);
          if (lst[a] == row)
            v += V(1.0);
          Ts[(warp * 8 + g) * ST + a] = v;
        }
    }
    // ---- second product: tile[:, chunk] += (-T) * U'[:, chunk]
    {
      V th[HOIST ? 2 : 1][HOIST ? KS2 : 1];
      for (int q = 0; q < nch; ++q, ++s)
      {
        step(s); // (its barrier also publishes -T)
        const V* buf = ring + (size_t)(s % STAGES) * STAGE_ELEMS;
        if constexpr (HOIST)
        {
          if (q == 0)
          {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int ks = 0; ks < KS2; ++ks)
                th[i][ks] = lds_v(Ts + ((2 * rp + i) * 8 + g) * ST + ks * 4 + t);
          }
        }
        const int j0 = q * KC;
        Acc<V> acc[2][JT];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < JT; ++j)
          {
            acc[i][j].set(0, cnext[i][j][0]);
            acc[i][j].set(1, cnext[i][j][1]);
          }
        if (q + 1 < nch)
          load_c(m0, j0 + KC);
        else if (it + 1 < niter)
          load_a(m0 + (int)gridDim.x * RT, 0);
#pragma unroll
        for (int ks = 0; ks < KS2; ++ks)
        {
          V a[2], b[JT];
#pragma unroll
          for (int i = 0; i < 2; ++i)
          {
            if constexpr (HOIST)
              a[i] = th[i][ks];
            else
              a[i] = lds_v(Ts + ((2 * rp + i) * 8 + g) * ST + ks * 4 + t);
          }
#pragma unroll
          for (int j = 0; j < JT; ++j)
            b[j] = lds_v(buf + (ks * 4 + t) * SP + (cp * JT + j) * 8 + g);
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j)
              acc[i][j].mma(a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < JT; ++j)
          {
            const int row = m0 + (2 * rp + i) * 8 + g, col = j0 + (cp * JT + j) * 8 + 2 * t;
            V* p          = Ainv + (size_t)row * lda + col;
            if (row < n && col + 1 < n)
            {
              if constexpr (C::CPLX)
              {
                const cx<double> v0 = acc[i][j].get(0), v1 = acc[i][j].get(1);
                st_hint2(reinterpret_cast<double*>(p), v0.re, v0.im, pol_drop);
                st_hint2(reinterpret_cast<double*>(p + 1), v1.re, v1.im, pol_drop);
              }
              else
                st_hint2(p, acc[i][j].get(0), acc[i][j].get(1), pol_drop);
            }
            else if (row < n && col < n)
              p[0] = acc[i][j].get(0);
          }
      }
    }
  }
  cp_async_wait<0>();
}
} // namespace wb64
#endif
} // namespace qmcb
