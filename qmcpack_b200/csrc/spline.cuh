// qmcpack_b200/csrc/spline.cuh -- multi-walker tricubic B-spline SPO evaluation for sm_100a.
//
// Replaces (reference paths under /root/reference/src):
//   SplineR2R::mw_evaluateVGLandDetRatioGrads   BsplineFactory/SplineR2R.cpp:414-582        (K1 in SURVEY 2.1)
//   SplineC2COMPTarget::mw_evaluateVGLandDetRatioGrads / evaluateVGLMultiPos  SplineC2COMPTarget.cpp:505-800 (K2,K3)
//   mw_evaluateDetRatios / mw_evaluateValue     SplineR2R.cpp:204-335, SPOSet.cpp:128-137   (K4)
// Arithmetic specification: spline2/MultiBsplineVGLH.hpp:120-227 (accumulation order i outer, j inner,
// prefactor products formed first), MultiBsplineEval_helper.hpp:32-93 (location in double, residual in ST),
// SplineR2R.cpp:338-374 / SplineC2C.cpp:200-277 (lattice contraction, sign, twist phase).
//
// Design (not a port of the OpenMP-target regions):
//   * one persistent CTA per SM slot; work unit = (walker, orbital tile of TILE real components)
//   * the table is described to the TMA unit as a 4-D tensor (component, z, y, x); one elected producer thread stages
//     a unit's whole 4x4x4 coefficient stencil (64 rows x TILE components, 48-64 KB) into shared memory with ONE
//     cp.async.bulk.tensor.4d request (SASS UTMALDG) completing on an mbarrier; STAGES units in flight per CTA, so HBM
//     latency is covered by the TMA engine instead of by registers/occupancy.  Out-of-range components of the last
//     tile are zero-filled by the TMA unit.
//   * consumer threads own VEC consecutive components each, read the staged rows conflict-free, keep
//     v/g(3)/h(6) in registers (the reference round-trips an 11-field scratch through global memory,
//     SplineR2R.cpp:459,485-498), apply the lattice contraction, write phi_vgl[5][nw][n] coalesced and
//     reduce ratio/grad per warp with shuffles; every (tile, warp) deposits its partial sums in a fixed slot
//     rg_partial[walker][part][4] and whoever consumes the ratio adds the parts in index order (sum_rg_parts): the
//     result is deterministic, needs no floating-point atomics, and the consumer warps never meet at a CTA barrier,
//     so one warp's epilogue overlaps the others' arithmetic
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace qmcb
{
enum SplineMode
{
  MODE_V       = 0, // values only (mw_evaluateValue / det ratios)
  MODE_VGL     = 1, // values, gradients, laplacians
};

template<typename ST>
struct SplineDev
{
  const ST* coefs;
  int M[3];
  int n_orb;  // orbitals requested (complex count for C2C)
  int n_spl;  // real spline components actually stored
  int npad;   // z stride
  long long xs, ys;
  double delta_inv[3];
  ST G[9];      // prim_lattice.G row-major
  ST symGG[6];  // GGt0, GGt1+GGt3, GGt2+GGt6, GGt4, GGt5+GGt7, GGt8
  int halfG[3];
  const ST* kcart; // [3][n_orb] SoA, C2C only
  const ST* mKK;   // [n_orb]
};

// arguments of one batched evaluation
template<typename ST, typename RT>
struct SplineArgs
{
  int nw;              // number of positions
  const RT* r;         // [nw][3] Cartesian
  int pdl_early;       // signal programmatic dependents at kernel start
  int l2_evict_first;  // stencil requests carry an evict_first L2 policy (env QMCB_SPL_EVICT)
  const ST* invrow;    // [n_rows][ld_inv] (VT == ST; complex interleaved for C2C) or nullptr
  const int* ref;      // optional [nw] row index into invrow (virtual-particle ratios); nullptr -> iw
  long long ld_inv;    // in VT elements
  ST* phi_vgl;         // MODE_VGL: [5][nw][n_orb] ; MODE_V: [nw][n_orb] ; may be nullptr
  ST* rg_partial;      // [nw][nparts][4] (x2 for complex): per-(tile, warp) partial ratio, gx, gy, gz (undivided); may be nullptr
  int nparts;          // ntiles * consumer warps
};

// sum the partial ratio/gradient dots of one walker in index order (NRED = 4 real, 8 complex)
template<typename ST, int NRED>
__host__ __device__ inline void sum_rg_parts(const ST* rg_partial, int iw, int nparts, ST out[NRED])
{
  for (int e = 0; e < NRED; ++e)
    out[e] = ST(0);
  const ST* p = rg_partial + (size_t)iw * nparts * NRED;
  for (int q = 0; q < nparts; ++q)
    for (int e = 0; e < NRED; ++e)
      out[e] += p[q * NRED + e];
}

#ifdef __CUDACC__
namespace ptx
{
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
  unsigned ok;
  const uint32_t a = smem_u32(bar);
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!ok);
}
// TMA bulk engine, 1-D: global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA tensor engine: one 4-D box (TILE components x 4 x 4 x 4) global -> shared (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
               "%5, %6}], [%2];" ::"r"(smem_u32(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// the same request with an L2 eviction policy: the stencil stream is read once per evaluation, the walkers' delayed-update
// rows (U, V, Binv, inverse rows) are read again by the next boundary kernel -- evict_first on the stream keeps them in L2
__device__ __forceinline__ void tma_load_4d_hint(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                                 uint64_t* bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, "
               "{%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* tmap)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
} // namespace ptx

// ref: Numerics/SplineBound.hpp:37-62 on the double-typed scaled coordinate, residual cast to ST
template<typename ST>
__device__ __forceinline__ void get_spline_bound(double x, int nmax, int& ind, ST& dx)
{
  if (x < 0)
  {
    ind = 0;
    dx  = ST(0);
  }
  else
  {
    double ipart;
    double fr = modf(x, &ipart);
    ind       = (int)ipart;
    dx        = (ST)fr;
    if (ind > nmax)
    {
      ind = nmax;
      dx  = (ST)(1.0 - 2.220446049250313e-16);
    }
  }
}

// ref: spline2/MultiBsplineData.hpp:30-87
template<typename T>
__device__ __forceinline__ void prefactors(T a[4], T da[4], T d2a[4], T tx)
{
  a[0]   = ((T(-1.0 / 6.0) * tx + T(3.0 / 6.0)) * tx + T(-3.0 / 6.0)) * tx + T(1.0 / 6.0);
  a[1]   = ((T(3.0 / 6.0) * tx + T(-6.0 / 6.0)) * tx + T(0.0)) * tx + T(4.0 / 6.0);
  a[2]   = ((T(-3.0 / 6.0) * tx + T(3.0 / 6.0)) * tx + T(3.0 / 6.0)) * tx + T(1.0 / 6.0);
  a[3]   = ((T(1.0 / 6.0) * tx + T(0.0)) * tx + T(0.0)) * tx + T(0.0);
  da[0]  = (T(-0.5) * tx + T(1.0)) * tx + T(-0.5);
  da[1]  = (T(1.5) * tx + T(-2.0)) * tx + T(0.0);
  da[2]  = (T(-1.5) * tx + T(1.0)) * tx + T(0.5);
  da[3]  = (T(0.5) * tx + T(0.0)) * tx + T(0.0);
  d2a[0] = T(-1.0) * tx + T(1.0);
  d2a[1] = T(3.0) * tx + T(-2.0);
  d2a[2] = T(-3.0) * tx + T(1.0);
  d2a[3] = T(1.0) * tx + T(0.0);
}

// Cartesian -> unit cell [0,1)^3 and bc sign.  ref: SplineR2R.h:156-170 (convertPos), CrystalLattice.h:187-198
template<typename ST, typename RT>
__device__ __forceinline__ int convert_pos(const SplineDev<ST>& S, const RT* r, ST ru[3])
{
  int bc_sign = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j)
  {
    ST v = ST(0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      v += ST(r[i]) * S.G[i * 3 + j];
    ru[j] = v;
  }
  const ST eps = sizeof(ST) == 4 ? ST(1.1920929e-07f) : ST(2.220446049250313e-16);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (-eps < ru[i] && ru[i] < 0)
      ru[i] = ST(0);
    else
    {
      ST img = floor(ru[i]);
      ru[i] -= img;
      bc_sign += S.halfG[i] * (int)img;
    }
  return bc_sign;
}

template<typename ST>
__device__ __forceinline__ void locate(const SplineDev<ST>& S, const ST ru[3], int ind[3], ST t[3])
{
#pragma unroll
  for (int d = 0; d < 3; ++d)
    get_spline_bound<ST>((double)ru[d] * S.delta_inv[d], S.M[d] - 1, ind[d], t[d]);
}

// per-stage unit header written by the producer warp (in ST): pre[16][8] = for every (i, j) the prefactor products
// {d2a_i b_j, da_i b_j, da_i db_j, a_i db_j, a_i d2b_j, a_i b_j, 0, 0}, then c[4], dc[4], d2c[4], then the bc sign (int)
constexpr int SPL_HDR     = 16 * 8 + 16;
constexpr int SPL_HDR_C   = 16 * 8;      // offset of c, dc, d2c
constexpr int SPL_HDR_SGN = 16 * 8 + 12; // offset of the sign word
constexpr int SPL_SCRATCH = 36;          // producer scratch: a, da, d2a, b, db, d2b, c, dc, d2c

template<typename ST, int TILE, int STAGES, int VEC>
struct SplineSmem
{
  static constexpr int ROWS        = 64;
  static constexpr size_t STAGE_B  = (size_t)ROWS * TILE * sizeof(ST);
  static constexpr size_t BAR_OFF  = STAGES * STAGE_B;
  static constexpr size_t HDR_OFF  = BAR_OFF + ((2 * STAGES * sizeof(uint64_t) + 15) / 16) * 16;
  static constexpr size_t BYTES    = HDR_OFF + (STAGES * SPL_HDR + SPL_SCRATCH + 4) * sizeof(ST) + 64;
};

template<typename ST>
__device__ __forceinline__ void load4(const ST* p, ST (&o)[4]);
template<>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4])
{
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
}
template<>
__device__ __forceinline__ void load4<double>(const double* p, double (&o)[4])
{
  const double2 v0 = reinterpret_cast<const double2*>(p)[0], v1 = reinterpret_cast<const double2*>(p)[1];
  o[0] = v0.x, o[1] = v0.y, o[2] = v1.x, o[3] = v1.y;
}

// MODE: SplineMode.  C2C: complex orbitals from pairs of components (VEC == 2), otherwise VEC == 1.
template<typename ST, typename RT, int TILE, int STAGES, int VEC, int MODE, bool C2C, int MINB>
__global__ void __launch_bounds__(TILE / VEC + 32, MINB)
    spline_gather_kernel(const __grid_constant__ CUtensorMap tmap, const SplineDev<ST> S, const SplineArgs<ST, RT> A,
                         const int ntiles)
{
  constexpr int NCONS = TILE / VEC; // consumer threads
  constexpr int ROWS  = 64;
  constexpr int NRED  = C2C ? 8 : 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ST* stage_base     = reinterpret_cast<ST*>(smem_raw);
  using SM = SplineSmem<ST, TILE, STAGES, VEC>;
  uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem_raw + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  ST* hdr_base        = reinterpret_cast<ST*>(smem_raw + SM::HDR_OFF);

  const int tid       = threadIdx.x;
  const int nunits    = A.nw * ntiles;
  const bool producer = tid >= NCONS;

  if (tid == 0)
  {
    for (int s = 0; s < STAGES; ++s)
    {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], NCONS / 32);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; the
  // positions and inverse rows it produced are read only from here on
  if (A.pdl_early)
    pdl_launch_dependents();
  pdl_wait();

  if (producer)
  {
    // ===== producer warp: everything that is uniform over the unit -- cell location, the 36 basis values and the 96
    // prefactor products -- is computed ONCE here (spread over the lanes) instead of by every consumer thread, parked
    // in the stage's header, and the stencil is requested with ONE tensor-TMA request issued by lane 0.  The
    // arithmetic is the consumers' former prologue verbatim, so results are bit-identical.
    const int pl = tid - NCONS;
    ST* scratch  = hdr_base + STAGES * SPL_HDR;
    if (pl == 0)
      ptx::prefetch_tensormap(&tmap);
    int q = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++q)
    {
      const int stage   = q % STAGES;
      const unsigned ph = (q / STAGES) & 1;
      const int iw = u / ntiles, tile = u - iw * ntiles;
      ST ru[3];
      const int bc_sign = convert_pos<ST, RT>(S, A.r + 3 * (size_t)iw, ru);
      int my_ind        = 0;
      if (pl < 3)
      {
        const ST rud     = pl == 0 ? ru[0] : (pl == 1 ? ru[1] : ru[2]);
        const double di  = pl == 0 ? S.delta_inv[0] : (pl == 1 ? S.delta_inv[1] : S.delta_inv[2]);
        const int nmax   = (pl == 0 ? S.M[0] : (pl == 1 ? S.M[1] : S.M[2])) - 1;
        ST t, p0[4], p1[4], p2[4];
        get_spline_bound<ST>((double)rud * di, nmax, my_ind, t);
        prefactors(p0, p1, p2, t);
#pragma unroll
        for (int e = 0; e < 4; ++e)
        {
          scratch[pl * 12 + e]     = p0[e];
          scratch[pl * 12 + 4 + e] = p1[e];
          scratch[pl * 12 + 8 + e] = p2[e];
        }
      }
      __syncwarp();
      const int i0 = __shfl_sync(0xffffffffu, my_ind, 0), i1 = __shfl_sync(0xffffffffu, my_ind, 1),
                i2 = __shfl_sync(0xffffffffu, my_ind, 2);
      ST o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        o[e] = ST(0);
      if (pl < 16)
      {
        const int i = pl >> 2, j = pl & 3;
        const ST ai = scratch[i], dai = scratch[4 + i], d2ai = scratch[8 + i];
        const ST bj = scratch[12 + j], dbj = scratch[16 + j], d2bj = scratch[20 + j];
        o[0] = d2ai * bj; // pre20
        o[1] = dai * bj;  // pre10
        o[2] = dai * dbj; // pre11
        o[3] = ai * dbj;  // pre01
        o[4] = ai * d2bj; // pre02
        o[5] = ai * bj;   // pre00
      }
      else if (pl < 19)
      {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = scratch[24 + (pl - 16) * 4 + e];
      }
      ptx::mbar_wait(&empty_bar[stage], ph ^ 1u);
      ST* hd = hdr_base + stage * SPL_HDR;
      if (pl < 16)
      {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          hd[pl * 8 + e] = o[e];
      }
      else if (pl < 19)
      {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          hd[SPL_HDR_C + (pl - 16) * 4 + e] = o[e];
      }
      else if (pl == 19)
        *reinterpret_cast<int*>(hd + SPL_HDR_SGN) = bc_sign;
      __syncwarp();
      if (pl == 0)
      {
        ptx::mbar_arrive_expect_tx(&full_bar[stage], (unsigned)(ROWS * TILE * sizeof(ST)));
        if (A.l2_evict_first)
          ptx::tma_load_4d_hint(stage_base + (size_t)stage * ROWS * TILE, &tmap, tile * TILE, i2, i1, i0, &full_bar[stage],
                                ptx::policy_evict_first());
        else
          ptx::tma_load_4d(stage_base + (size_t)stage * ROWS * TILE, &tmap, tile * TILE, i2, i1, i0, &full_bar[stage]);
      }
      __syncwarp();
    }
    return;
  }

  // ===== consumers =====
  const int warp = tid >> 5, lane = tid & 31;
  int q = 0;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++q)
  {
    const int stage = q % STAGES;
    const unsigned ph = (q / STAGES) & 1;
    const int iw = u / ntiles, tile = u - iw * ntiles;
    const RT* rpos = A.r + 3 * (size_t)iw;

    ST v[VEC], gx[VEC], gy[VEC], gz[VEC], hxx[VEC], hxy[VEC], hxz[VEC], hyy[VEC], hyz[VEC], hzz[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      v[e] = gx[e] = gy[e] = gz[e] = hxx[e] = hxy[e] = hxz[e] = hyy[e] = hyz[e] = hzz[e] = ST(0);

    // fetch this thread's inverse-row element(s) while the stencil is still in flight
    const int row_id = A.ref ? A.ref[iw] : iw;
    ST wr = ST(0), wi = ST(0);
    if (A.invrow)
    {
      if (!C2C)
      {
        const int m = tile * TILE + tid;
        if (m < S.n_orb)
          wr = A.invrow[(size_t)row_id * A.ld_inv + m];
      }
      else
      {
        const int jorb = (tile * TILE) / 2 + tid;
        if (jorb < S.n_orb)
        {
          wr = A.invrow[((size_t)row_id * A.ld_inv + jorb) * 2];
          wi = A.invrow[((size_t)row_id * A.ld_inv + jorb) * 2 + 1];
        }
      }
    }
    ptx::mbar_wait(&full_bar[stage], ph);
    const ST* sm = stage_base + (size_t)stage * ROWS * TILE + tid * VEC;
    const ST* hd = hdr_base + stage * SPL_HDR;
    ST c[4], dc[4], d2c[4];
    load4<ST>(hd + SPL_HDR_C, c);
    if (MODE != MODE_V)
    {
      load4<ST>(hd + SPL_HDR_C + 4, dc);
      load4<ST>(hd + SPL_HDR_C + 8, d2c);
    }
    const int bc_sign = *reinterpret_cast<const int*>(hd + SPL_HDR_SGN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        const ST* p0 = sm + (size_t)((i * 4 + j) * 4) * TILE;
        ST pa[4], pb[4];
        load4<ST>(hd + (i * 4 + j) * 8 + 4, pb);
        const ST pre00 = pb[1];
        if (MODE == MODE_V)
        {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            v[e] += pre00 * (c[0] * p0[e] + c[1] * p0[TILE + e] + c[2] * p0[2 * TILE + e] + c[3] * p0[3 * TILE + e]);
        }
        else
        {
          load4<ST>(hd + (i * 4 + j) * 8, pa);
          const ST pre20 = pa[0], pre10 = pa[1], pre11 = pa[2], pre01 = pa[3], pre02 = pb[0];
#pragma unroll
          for (int e = 0; e < VEC; ++e)
          {
            const ST c0 = p0[e], c1 = p0[TILE + e], c2 = p0[2 * TILE + e], c3 = p0[3 * TILE + e];
            const ST sum0 = c[0] * c0 + c[1] * c1 + c[2] * c2 + c[3] * c3;
            const ST sum1 = dc[0] * c0 + dc[1] * c1 + dc[2] * c2 + dc[3] * c3;
            const ST sum2 = d2c[0] * c0 + d2c[1] * c1 + d2c[2] * c2 + d2c[3] * c3;
            hxx[e] += pre20 * sum0;
            hxy[e] += pre11 * sum0;
            hxz[e] += pre10 * sum1;
            hyy[e] += pre02 * sum0;
            hyz[e] += pre01 * sum1;
            hzz[e] += pre00 * sum2;
            gx[e] += pre10 * sum0;
            gy[e] += pre01 * sum0;
            gz[e] += pre00 * sum1;
            v[e] += pre00 * sum0;
          }
        }
      }
    // stage consumed: one arrival per consumer warp
    __syncwarp();
    if (lane == 0)
      ptx::mbar_arrive(&empty_bar[stage]);

    // ---- epilogue: lattice units -> Cartesian, sign / twist phase, outputs, dot with the inverse row
    const ST dxInv = (ST)S.delta_inv[0], dyInv = (ST)S.delta_inv[1], dzInv = (ST)S.delta_inv[2];
    ST acc[NRED];
#pragma unroll
    for (int e = 0; e < NRED; ++e)
      acc[e] = ST(0);
    const size_t fstride = (size_t)A.nw * S.n_orb * (C2C ? 2 : 1);

    if (!C2C)
    {
      const int m = tile * TILE + tid; // orbital
      if (m < S.n_orb)
      {
        const ST sgn = (bc_sign & 1) ? ST(-1) : ST(1);
        if (MODE == MODE_V)
        {
          const ST psi = sgn * v[0];
          if (A.phi_vgl)
            A.phi_vgl[(size_t)iw * S.n_orb + m] = psi;
          acc[0] = psi * wr;
        }
        else
        {
          const ST g0 = gx[0] * dxInv, g1 = gy[0] * dyInv, g2 = gz[0] * dzInv;
          const ST h00 = hxx[0] * (dxInv * dxInv), h11 = hyy[0] * (dyInv * dyInv), h22 = hzz[0] * (dzInv * dzInv);
          const ST h01 = hxy[0] * (dxInv * dyInv), h02 = hxz[0] * (dxInv * dzInv), h12 = hyz[0] * (dyInv * dzInv);
          const ST psi = sgn * v[0];
          const ST dx  = sgn * (S.G[0] * g0 + S.G[1] * g1 + S.G[2] * g2);
          const ST dy  = sgn * (S.G[3] * g0 + S.G[4] * g1 + S.G[5] * g2);
          const ST dz  = sgn * (S.G[6] * g0 + S.G[7] * g1 + S.G[8] * g2);
          const ST lap = sgn *
              (h00 * S.symGG[0] + h01 * S.symGG[1] + h02 * S.symGG[2] + h11 * S.symGG[3] + h12 * S.symGG[4] +
               h22 * S.symGG[5]);
          if (A.phi_vgl)
          {
            ST* out = A.phi_vgl + (size_t)iw * S.n_orb + m;
            out[0]           = psi;
            out[fstride]     = dx;
            out[2 * fstride] = dy;
            out[3 * fstride] = dz;
            out[4 * fstride] = lap;
          }
          acc[0] = psi * wr;
          acc[1] = dx * wr;
          acc[2] = dy * wr;
          acc[3] = dz * wr;
        }
      }
    }
    else
    {
      const int jorb = (tile * TILE) / 2 + tid; // complex orbital; components (2j, 2j+1) = v[0], v[1]
      if (jorb < S.n_orb)
      {
        const ST kX = S.kcart[jorb], kY = S.kcart[S.n_orb + jorb], kZ = S.kcart[2 * S.n_orb + jorb];
        const ST x = (ST)rpos[0], y = (ST)rpos[1], z = (ST)rpos[2];
        ST s, cs;
        sincos(-(x * kX + y * kY + z * kZ), &s, &cs);
        const ST val_r = v[0], val_i = v[1];
        const ST psi_r = cs * val_r - s * val_i, psi_i = cs * val_i + s * val_r;
        if (MODE == MODE_V)
        {
          if (A.phi_vgl)
          {
            ST* out = A.phi_vgl + ((size_t)iw * S.n_orb + jorb) * 2;
            out[0]  = psi_r;
            out[1]  = psi_i;
          }
          acc[0] = psi_r * wr - psi_i * wi;
          acc[1] = psi_r * wi + psi_i * wr;
        }
        else
        {
          const ST g0r = gx[0] * dxInv, g1r = gy[0] * dyInv, g2r = gz[0] * dzInv;
          const ST g0i = gx[1] * dxInv, g1i = gy[1] * dyInv, g2i = gz[1] * dzInv;
          const ST dX_r = S.G[0] * g0r + S.G[1] * g1r + S.G[2] * g2r;
          const ST dY_r = S.G[3] * g0r + S.G[4] * g1r + S.G[5] * g2r;
          const ST dZ_r = S.G[6] * g0r + S.G[7] * g1r + S.G[8] * g2r;
          const ST dX_i = S.G[0] * g0i + S.G[1] * g1i + S.G[2] * g2i;
          const ST dY_i = S.G[3] * g0i + S.G[4] * g1i + S.G[5] * g2i;
          const ST dZ_i = S.G[6] * g0i + S.G[7] * g1i + S.G[8] * g2i;
          const ST gX_r = dX_r + val_i * kX, gY_r = dY_r + val_i * kY, gZ_r = dZ_r + val_i * kZ;
          const ST gX_i = dX_i - val_r * kX, gY_i = dY_i - val_r * kY, gZ_i = dZ_i - val_r * kZ;
          const ST sxx = dxInv * dxInv, syy = dyInv * dyInv, szz = dzInv * dzInv, sxy = dxInv * dyInv,
                   sxz = dxInv * dzInv, syz = dyInv * dzInv;
          const ST lcart_r = hxx[0] * sxx * S.symGG[0] + hxy[0] * sxy * S.symGG[1] + hxz[0] * sxz * S.symGG[2] +
              hyy[0] * syy * S.symGG[3] + hyz[0] * syz * S.symGG[4] + hzz[0] * szz * S.symGG[5];
          const ST lcart_i = hxx[1] * sxx * S.symGG[0] + hxy[1] * sxy * S.symGG[1] + hxz[1] * sxz * S.symGG[2] +
              hyy[1] * syy * S.symGG[3] + hyz[1] * syz * S.symGG[4] + hzz[1] * szz * S.symGG[5];
          const ST mKK   = S.mKK[jorb];
          const ST lap_r = lcart_r + mKK * val_r + ST(2) * (kX * dX_i + kY * dY_i + kZ * dZ_i);
          const ST lap_i = lcart_i + mKK * val_i - ST(2) * (kX * dX_r + kY * dY_r + kZ * dZ_r);
          const ST o[10] = {psi_r,
                            psi_i,
                            cs * gX_r - s * gX_i,
                            cs * gX_i + s * gX_r,
                            cs * gY_r - s * gY_i,
                            cs * gY_i + s * gY_r,
                            cs * gZ_r - s * gZ_i,
                            cs * gZ_i + s * gZ_r,
                            cs * lap_r - s * lap_i,
                            cs * lap_i + s * lap_r};
          if (A.phi_vgl)
          {
            ST* out = A.phi_vgl + ((size_t)iw * S.n_orb + jorb) * 2;
#pragma unroll
            for (int f = 0; f < 5; ++f)
            {
              out[f * fstride]     = o[2 * f];
              out[f * fstride + 1] = o[2 * f + 1];
            }
          }
#pragma unroll
          for (int f = 0; f < 4; ++f)
          {
            acc[2 * f]     = o[2 * f] * wr - o[2 * f + 1] * wi;
            acc[2 * f + 1] = o[2 * f] * wi + o[2 * f + 1] * wr;
          }
        }
      }
    }

    if (A.rg_partial)
    {
      // per-warp partial sums in a fixed slot; consumers of the ratio add the slots in index order
#pragma unroll
      for (int e = 0; e < NRED; ++e)
        acc[e] = warp_sum(acc[e]);
      if (lane < NRED)
      {
        ST mine = acc[0];
#pragma unroll
        for (int e = 1; e < NRED; ++e)
          mine = (lane == e) ? acc[e] : mine;
        A.rg_partial[((size_t)iw * A.nparts + tile * (NCONS / 32) + warp) * NRED + lane] = mine;
      }
    }
  }
}

// rg[nw][NRED] = sum of the parts (standalone API and tests; inside the sweep the consumers add the parts themselves)
template<typename ST, int NRED>
__global__ void spline_finalize_rg_kernel(const ST* rg_partial, int nw, int nparts, ST* rg)
{
  const int iw = blockIdx.x * blockDim.x + threadIdx.x;
  if (iw >= nw)
    return;
  ST out[NRED];
  sum_rg_parts<ST, NRED>(rg_partial, iw, nparts, out);
  for (int e = 0; e < NRED; ++e)
    rg[(size_t)iw * NRED + e] = out[e];
}
#endif // __CUDACC__

} // namespace qmcb
