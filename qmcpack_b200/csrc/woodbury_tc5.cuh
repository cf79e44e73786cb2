// qmcpack_b200/csrc/woodbury_tc5.cuh -- rank-k Woodbury flush on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contraction as woodbury.cuh (DelayedUpdateBatched::mw_updateInvMat, Fermion/DelayedUpdateBatched.h:675-738):
//     T[m][a]  = sum_j Ainv[m][j] U[a][j]  - (m == list[a])
//     U'[a][j] = sum_b Binv[a][b] V[b][j]
//     Ainv[m][j] -= sum_a T[m][a] U'[a][j]
// One CTA owns a 128-row tile of one walker's Ainv; the accumulators live in tensor memory:
//     D1 [128 lanes x 32 columns]  = tile * U^T            (K = n, streamed in 32-column chunks, double buffered)
//     D2 [128 lanes x n columns]   = (-T) * U'              (K = 32)
// Operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout (rows of 32 floats = 128 bytes, the
// 16-byte chunk index XOR-ed with row % 8, 8-row groups 1024 bytes apart) by the CUDA cores, which also split every value
// x = hi + lo (hi = x with the 13 low mantissa bits cleared, lo = x - hi, exact): three tcgen05.mma.kind::tf32 products
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo accumulate into the same TMEM tile, which gives the update FP32-level accuracy on the
// TF32 datapath (a single TF32 product would destroy the inverse).  One elected thread issues the MMAs; completion comes
// back through tcgen05.commit on mbarriers; the epilogue reads TMEM with tcgen05.ld, transposes through shared memory and
// adds the tile (re-read from L2) with coalesced 128-byte rows.  HBM traffic: one read + one write of Ainv.
#pragma once
#include "common.cuh"
#include "det.cuh"

namespace qmcb
{
#ifdef __CUDACC__
namespace wb5
{
constexpr int TM  = 128; // Ainv rows per CTA = MMA M = TMEM lanes
constexpr int KC  = 32;  // K chunk = one 128-byte swizzle row of floats
constexpr int KD  = 32;  // delay slots handled (c <= KD)
constexpr int TPB = 256;

// dynamic shared memory: two stages of {A_hi, A_lo [128][32], U_hi, U_lo [32][32]}; +1024 for manual alignment
constexpr int STAGE_FLOATS = 2 * TM * KC + 2 * KD * KC;
inline size_t smem_bytes(int) { return (size_t)2 * STAGE_FLOATS * 4 + 1024; }
// n must be a multiple of 32 (K chunks) and the accumulators must fit the 512 TMEM columns
// n: multiple of 64 (K chunks of 32; each half of the second product's N in 32-column epilogue pieces); the accumulators
// (32 + n columns) must fit the 512 TMEM columns
inline bool eligible(int n, int k, int c) { return n % 64 == 0 && n >= 64 && k <= KD && c <= KD; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "WAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\t"
               "bra WAIT_%=;\n\t"
               "DONE_%=:\n\t"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor fields): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte offset
// (8-row group pitch = 1024 B) >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr)
{
  const uint64_t addr = (uint64_t)((smem_u32(smem_ptr) & 0x3FFFFu) >> 4);
  return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N)
{
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
               "}\n" ::"r"(tmem_d),
               "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 32 lanes (this warp's quarter of the 128) x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i)
    v[i] = __uint_as_float(r[i]);
}

// float offset of element (row, k) inside a K-major SWIZZLE_128B tile whose rows are KC = 32 floats (128 bytes)
__device__ __forceinline__ int sw_off(int row, int k) { return row * KC + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3)); }

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo)
{
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  lo   = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

// grid = (ceil(n / 128), nw); c <= 32 pending delays; n % 64 == 0.  Two CTAs per SM (85 KB of shared memory, 256 of the
// 512 TMEM columns and <= 128 registers each): the phases of one CTA are chains of memory latencies, the second CTA on
// the SM fills them.
// up_ready: D.Up already holds U' = Binv V (left there by the walker-segment kernel, segment.cuh)
__global__ void __launch_bounds__(TPB, 2) woodbury_flush_tc5_kernel(const DetDev<float> D, const int c, const int up_ready)
{
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(16) float BinvT[KD][KD + 4]; // BinvT[b][a] = Binv[a][b] (float4 reads over a)
  __shared__ int lst[KD];
  __shared__ __align__(8) uint64_t bar_stage[2];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;

  const int n = D.n, lda = D.lda, k = D.k;
  const int iw = blockIdx.y, m0 = blockIdx.x * TM; // tile index fastest: a walker's CTAs run together and share U, V in L2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = n / KC;
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 bytes).  One stage = A_hi, A_lo [128][32] + U_hi, U_lo [32][32]
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  float* S0 = reinterpret_cast<float*>(base);
  auto Aptr = [&](int stage, int part) { return S0 + (size_t)stage * STAGE_FLOATS + (size_t)part * TM * KC; };
  auto Uptr = [&](int stage, int part) { return S0 + (size_t)stage * STAGE_FLOATS + (size_t)2 * TM * KC + (size_t)part * KD * KC; };

  const float* U = D.U + (size_t)iw * k * n;
  const float* V = D.V + (size_t)iw * k * n;
  const float* B = D.Binv + (size_t)iw * k * k;
  float* Ainv    = D.Ainv + (size_t)iw * n * lda;

  // ---- operand loaders.  A chunk [128 x 32]: thread -> rows tid/8 + 32 i (i < 4), 16-byte piece tid % 8 of the 128-byte
  //      row; U chunk [32 x 32]: thread -> row tid/8, piece tid % 8.  Two chunks stay in flight in registers.
  const int ar = tid >> 3, akq = tid & 7;
  float4 areg[2][4], ureg[2];
  auto load_chunk = [&](int kc, float4 (&dst)[4], float4& du) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const int row = m0 + ar + 32 * i;
      dst[i] = row < n ? *reinterpret_cast<const float4*>(Ainv + (size_t)row * lda + kc * KC + akq * 4)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    du = ar < c ? __ldg(reinterpret_cast<const float4*>(U + (size_t)ar * n + kc * KC + akq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_chunk = [&](int stage, const float4 (&src)[4], const float4& su) {
    float* hi_t = Aptr(stage, 0);
    float* lo_t = Aptr(stage, 1);
    float4 hi, lo;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const int r = ar + 32 * i;
      split4(src[i], hi, lo);
      const int off = r * KC + ((akq ^ (r & 7)) << 2);
      *reinterpret_cast<float4*>(hi_t + off) = hi;
      *reinterpret_cast<float4*>(lo_t + off) = lo;
    }
    split4(su, hi, lo);
    const int off = ar * KC + ((akq ^ (ar & 7)) << 2);
    *reinterpret_cast<float4*>(Uptr(stage, 0) + off) = hi;
    *reinterpret_cast<float4*>(Uptr(stage, 1) + off) = lo;
  };
  load_chunk(0, areg[0], ureg[0]);
  if (nchunks > 1)
    load_chunk(1, areg[1], ureg[1]);

  // ---- set-up: TMEM allocation (256 columns: D1 in [0,32), one 128-column piece of D2 in [32,160)), barriers
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32)
  {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Binv (transposed, zero beyond the c pending slots; only needed when U' is formed here) and the delay list
  if (!up_ready)
    for (int e = tid; e < KD * KD; e += TPB)
    {
      const int a = e / KD, b = e - a * KD;
      BinvT[b][a] = (a < c && b < c) ? __ldg(B + a * k + b) : 0.f;
    }
  if (tid < KD)
    lst[tid] = tid < c ? D.list[(size_t)iw * k + tid] : -1;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmemD1 = tmem, tmemD2 = tmem + KD; // column offsets

  // ---- GEMM 1: D1 = tile[128 x n] * U^T, K streamed in chunks of 32 (4 MMA k-steps of 8), 3 split products each
  const uint32_t idesc1 = make_idesc(TM, KD);
  for (int kc = 0; kc < nchunks; ++kc)
  {
    const int stage = kc & 1;
    if (kc >= 2)
      mbar_wait(&bar_stage[stage], (unsigned)(((kc >> 1) - 1) & 1)); // the MMAs that read this stage have retired
    if (stage == 0)
    {
      store_chunk(0, areg[0], ureg[0]);
      if (kc + 2 < nchunks)
        load_chunk(kc + 2, areg[0], ureg[0]);
    }
    else
    {
      store_chunk(1, areg[1], ureg[1]);
      if (kc + 2 < nchunks)
        load_chunk(kc + 2, areg[1], ureg[1]);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0)
    {
      fence_after_sync();
      const uint64_t a_hi = make_desc(Aptr(stage, 0)), a_lo = make_desc(Aptr(stage, 1));
      const uint64_t b_hi = make_desc(Uptr(stage, 0)), b_lo = make_desc(Uptr(stage, 1));
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks)
      {
        const uint64_t adv = (uint64_t)(ks * 2); // 8 floats = 32 bytes = 2 x 16-byte units inside the swizzle atom
        mma_tf32_ss(tmemD1, a_lo + adv, b_hi + adv, idesc1, (kc | ks) != 0);
        mma_tf32_ss(tmemD1, a_hi + adv, b_lo + adv, idesc1, 1u);
        mma_tf32_ss(tmemD1, a_hi + adv, b_hi + adv, idesc1, 1u);
      }
      mma_commit(&bar_stage[stage]);
      if (kc == nchunks - 1)
        mma_commit(&bar_done);
    }
  }
  unsigned done_parity = 0;
  mbar_wait(&bar_done, done_parity);
  done_parity ^= 1u;
  fence_after_sync();

  // ---- T epilogue (warps 0-3: one TMEM lane = one row each): -T with the list fix-up, split, stored as the A operand
  //      of the second product over stage 0's A tiles
  if (warp < 4)
  {
    float t[32];
    tmem_ld32(tmemD1 + ((uint32_t)(warp * 32) << 16), t);
    const int r = warp * 32 + lane;
    float* hi_t = Aptr(0, 0);
    float* lo_t = Aptr(0, 1);
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      float4 v;
      v.x = -t[4 * q] + (lst[4 * q] == m0 + r ? 1.f : 0.f);
      v.y = -t[4 * q + 1] + (lst[4 * q + 1] == m0 + r ? 1.f : 0.f);
      v.z = -t[4 * q + 2] + (lst[4 * q + 2] == m0 + r ? 1.f : 0.f);
      v.w = -t[4 * q + 3] + (lst[4 * q + 3] == m0 + r ? 1.f : 0.f);
      float4 hi, lo;
      split4(v, hi, lo);
      const int off = r * KC + ((q ^ (r & 7)) << 2);
      *reinterpret_cast<float4*>(hi_t + off) = hi;
      *reinterpret_cast<float4*>(lo_t + off) = lo;
    }
  }

  // ---- GEMM 2 in pieces of 128 columns: D2 = (-T)[128 x 32] * U'[32 x piece]; the U'^T piece (B operand, hi + lo) and
  //      later the epilogue's transpose patches live in stage 1
  float* Bp_hi = Aptr(1, 0);
  float* Bp_lo = Aptr(1, 1);
  float* patch = Aptr(1, 0) + (size_t)warp * 16 * 33; // 8 x 2.1 KB
  for (int j0 = 0; j0 < n; j0 += 128)
  {
    const int np = (n - j0) < 128 ? (n - j0) : 128; // columns in this piece (multiple of 64)
    // U'[a][j] = sum_b Binv[a][b] V[b][j]: thread -> column jj = tid % 128, slots 16 (tid / 128) .. +15
    {
      const int jj = tid & 127, a0 = (tid >> 7) * 16;
      if (jj < np && up_ready)
      {
        // U' comes from memory: sixteen coalesced loads per thread, split, stored as the B operand
        const float* Up = D.Up + (size_t)iw * k * n + j0 + jj;
        float acc[16];
#pragma unroll
        for (int a = 0; a < 16; ++a)
          acc[a] = a0 + a < c ? __ldg(Up + (size_t)(a0 + a) * n) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          float4 hi, lo;
          split4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), hi, lo);
          const int off = jj * KC + ((((a0 >> 2) + q) ^ (jj & 7)) << 2);
          *reinterpret_cast<float4*>(Bp_hi + off) = hi;
          *reinterpret_cast<float4*>(Bp_lo + off) = lo;
        }
      }
      else if (jj < np)
      {
        float vv[KD]; // the whole column of V in flight at once
#pragma unroll
        for (int b = 0; b < KD; ++b)
          vv[b] = b < c ? __ldg(V + (size_t)b * n + j0 + jj) : 0.f;
        float acc[16];
#pragma unroll
        for (int a = 0; a < 16; ++a)
          acc[a] = 0.f;
#pragma unroll
        for (int b = 0; b < KD; ++b)
        {
          const float v = vv[b];
#pragma unroll
          for (int q = 0; q < 4; ++q)
          {
            const float4 bb = *reinterpret_cast<const float4*>(&BinvT[b][a0 + 4 * q]);
            acc[4 * q] += bb.x * v;
            acc[4 * q + 1] += bb.y * v;
            acc[4 * q + 2] += bb.z * v;
            acc[4 * q + 3] += bb.w * v;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          float4 hi, lo;
          split4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), hi, lo);
          const int off = jj * KC + ((((a0 >> 2) + q) ^ (jj & 7)) << 2);
          *reinterpret_cast<float4*>(Bp_hi + off) = hi;
          *reinterpret_cast<float4*>(Bp_lo + off) = lo;
        }
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0)
    {
      fence_after_sync();
      const uint32_t idesc2 = make_idesc(TM, np);
      const uint64_t a_hi = make_desc(Aptr(0, 0)), a_lo = make_desc(Aptr(0, 1));
      const uint64_t b_hi = make_desc(Bp_hi), b_lo = make_desc(Bp_lo);
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks)
      {
        const uint64_t adv = (uint64_t)(ks * 2);
        mma_tf32_ss(tmemD2, a_lo + adv, b_hi + adv, idesc2, ks != 0);
        mma_tf32_ss(tmemD2, a_hi + adv, b_lo + adv, idesc2, 1u);
        mma_tf32_ss(tmemD2, a_hi + adv, b_hi + adv, idesc2, 1u);
      }
      mma_commit(&bar_done);
    }
    mbar_wait(&bar_done, done_parity);
    done_parity ^= 1u;
    fence_after_sync();

    // epilogue of the piece: Ainv tile += D2.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (its quarter) and the 32-column
    // slices w / 4, w / 4 + 2, ... of the piece, transposes 16 rows at a time through its padded patch (the U' piece is
    // dead) and updates the rows (L2 hits: this CTA streamed them a moment ago) with coalesced 128-byte accesses
    {
      const int q = warp & 3;
      for (int c0 = (warp >> 2) * 32; c0 < np; c0 += 64)
      {
        float v[32];
        tmem_ld32(tmemD2 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow)
        {
          if ((lane >> 4) == hrow)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              patch[(lane & 15) * 33 + i] = v[i];
          __syncwarp();
          float gv[16];
          float* g0 = Ainv + (size_t)(m0 + q * 32 + hrow * 16) * lda + j0 + c0 + lane;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            gv[rr] = (m0 + q * 32 + hrow * 16 + rr < n) ? __ldcg(g0 + (size_t)rr * lda) : 0.f;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            if (m0 + q * 32 + hrow * 16 + rr < n)
              __stcs(g0 + (size_t)rr * lda, gv[rr] + patch[rr * 33 + lane]);
          __syncwarp();
        }
      }
      // (requesting both halves' tile rows before tcgen05.ld was measured: 270 vs 250 us per flush in the sweep -- the
      // 32 extra live registers spill at the 128-register budget of two CTAs per SM)
    }
    fence_before_sync();
    __syncthreads(); // patches (and D2) are free for the next piece
    fence_after_sync();
  }
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
} // namespace wb5
#endif
} // namespace qmcb
