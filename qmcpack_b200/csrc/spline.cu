// qmcpack_b200/csrc/spline.cu -- host side of the spline SPOSet: table upload, launch configuration.
#include "internal.h"
#include <cstdlib>
#include "spline.cuh"
#include <cstring>

namespace qmcb
{
std::atomic<unsigned long long> g_launch_count{0};
static int pdl_mode_from_env()
{
  const char* e = std::getenv("QMCB_PDL");
  return e ? std::atoi(e) : 2;
}
int g_pdl_mode = pdl_mode_from_env();

namespace
{
int sm_count()
{
  int dev = 0, n = 0;
  QMCB_CUDA(cudaGetDevice(&dev));
  QMCB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

// tile/stage shapes per (storage type, kind); shared memory per CTA = STAGES * 64 * TILE * sizeof(ST)
template<typename ST, bool C2C>
struct Shape;
// MINB CTAs per SM share the 227 KB of shared memory and the register file
template<>
struct Shape<float, false>
{
#ifndef QMCB_SPL_TILE // tuning experiments: -DQMCB_SPL_TILE=.. -DQMCB_SPL_STAGES=.. -DQMCB_SPL_MINB=..
#define QMCB_SPL_TILE 192
#define QMCB_SPL_STAGES 2
#define QMCB_SPL_MINB 2
#endif
  static constexpr int TILE = QMCB_SPL_TILE, STAGES = QMCB_SPL_STAGES, VEC = 1, MINB = QMCB_SPL_MINB;
}; // 48 KB / stage, 2 CTAs/SM
template<>
struct Shape<double, false>
{
  static constexpr int TILE = 128, STAGES = 3, VEC = 1, MINB = 1;
}; // 64 KB / stage
template<>
struct Shape<float, true>
{
  static constexpr int TILE = 256, STAGES = 3, VEC = 2, MINB = 1;
}; // 64 KB / stage
template<>
struct Shape<double, true>
{
  static constexpr int TILE = 128, STAGES = 3, VEC = 2, MINB = 1;
}; // 64 KB / stage

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn)
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    QMCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// the table as a 4-D tensor (component fastest, then z, y, x); box = TILE components x 4 x 4 x 4 grid planes
template<typename ST>
CUtensorMap make_tensor_map(const ST* coefs, const int grid[3], size_t npad, int tile, int bz = 4, int by = 4, int bx = 4)
{
  CUtensorMap m;
  const cuuint64_t dims[4]    = {(cuuint64_t)npad, (cuuint64_t)(grid[2] + 3), (cuuint64_t)(grid[1] + 3),
                                 (cuuint64_t)(grid[0] + 3)};
  const cuuint64_t strides[3] = {(cuuint64_t)npad * sizeof(ST), (cuuint64_t)npad * sizeof(ST) * (grid[2] + 3),
                                 (cuuint64_t)npad * sizeof(ST) * (grid[2] + 3) * (grid[1] + 3)};
  const cuuint32_t box[4]     = {(cuuint32_t)tile, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx};
  const cuuint32_t estr[4]    = {1, 1, 1, 1};
  const CUresult r = encode_tiled()(&m, sizeof(ST) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4,
                                    const_cast<ST*>(coefs), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}
} // namespace

template<typename ST>
struct SplineSPO : SplineSPOBase
{
  DevBuf<ST> coefs, kc, mkk;
  SplineDev<ST> dev;
  CUtensorMap tmap; // box shape follows Shape<ST, kind == C2C>
  CUtensorMap tmap_seg; // 192 components x 4 z x 2 y x 1 x: one slab of the stencil (segment.cuh)
  size_t table_bytes() const override { return coefs.bytes(); }
  const void* dev_desc() const override { return &dev; }
  const void* seg_tensor_map() const override { return &tmap_seg; }

  SplineSPO(int prec, int kind_, const int g[3], int norb, int nspl, size_t npad_, const void* host, const double G_[9],
            const int hG[3], const double* kcart)
  {
    precision = prec;
    kind      = kind_;
    n_orb     = norb;
    n_spl     = nspl;
    npad      = npad_;
    for (int d = 0; d < 3; ++d)
    {
      grid[d]  = g[d];
      halfG[d] = hG ? hG[d] : 0;
    }
    std::memcpy(G, G_, sizeof(G));
    if (npad % (64 / sizeof(ST)) != 0)
      throw std::runtime_error("spline table: npad must be a multiple of 64 bytes (getAlignedSize)");
    if ((size_t)nspl > npad || (kind == QMCB_C2C ? 2 * norb : norb) > nspl)
      throw std::runtime_error("spline table: inconsistent n_orb / n_spl / npad");
    const size_t total = (size_t)(g[0] + 3) * (g[1] + 3) * (g[2] + 3) * npad;
    coefs.alloc(total, false);
    QMCB_CUDA(cudaMemcpy(coefs.p, host, total * sizeof(ST), cudaMemcpyHostToDevice));
    dev.coefs = coefs.p;
    dev.n_orb = norb;
    dev.n_spl = nspl;
    dev.npad  = (int)npad;
    dev.ys    = (long long)(g[2] + 3) * (long long)npad;
    dev.xs    = (long long)(g[1] + 3) * dev.ys;
    for (int d = 0; d < 3; ++d)
    {
      dev.M[d] = g[d];
      // MultiBsplineBase.hpp:95-125: delta = (end-start)/(N-3) = 1/M, delta_inv = 1/delta
      const double delta = 1.0 / (double)g[d];
      dev.delta_inv[d]   = 1.0 / delta;
      dev.halfG[d]       = halfG[d];
    }
    double GGt[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
      {
        double s = 0;
        for (int k = 0; k < 3; ++k)
          s += G[k * 3 + i] * G[k * 3 + j];
        GGt[i * 3 + j] = s;
      }
    for (int i = 0; i < 9; ++i)
      dev.G[i] = (ST)G[i];
    dev.symGG[0] = (ST)GGt[0];
    dev.symGG[1] = (ST)GGt[1] + (ST)GGt[3];
    dev.symGG[2] = (ST)GGt[2] + (ST)GGt[6];
    dev.symGG[3] = (ST)GGt[4];
    dev.symGG[4] = (ST)GGt[5] + (ST)GGt[7];
    dev.symGG[5] = (ST)GGt[8];
    dev.kcart    = nullptr;
    dev.mKK      = nullptr;
    tmap = make_tensor_map<ST>(coefs.p, grid, npad, kind == QMCB_C2C ? Shape<ST, true>::TILE : Shape<ST, false>::TILE);
    tmap_seg = make_tensor_map<ST>(coefs.p, grid, npad, 192, 4, 2, 1);
    if (kind == QMCB_C2C)
    {
      if (!kcart)
        throw std::runtime_error("SplineC2C needs the Cartesian twist vectors (kcart)");
      std::vector<ST> k(3 * (size_t)norb), m(norb);
      for (int j = 0; j < norb; ++j)
      {
        for (int d = 0; d < 3; ++d)
          k[(size_t)d * norb + j] = (ST)kcart[3 * j + d];
        m[j] = (ST)(-(kcart[3 * j] * kcart[3 * j] + kcart[3 * j + 1] * kcart[3 * j + 1] +
                      kcart[3 * j + 2] * kcart[3 * j + 2]));
      }
      kc.alloc(k.size(), false);
      mkk.alloc(m.size(), false);
      QMCB_CUDA(cudaMemcpy(kc.p, k.data(), k.size() * sizeof(ST), cudaMemcpyHostToDevice));
      QMCB_CUDA(cudaMemcpy(mkk.p, m.data(), m.size() * sizeof(ST), cudaMemcpyHostToDevice));
      dev.kcart = kc.p;
      dev.mKK   = mkk.p;
    }
  }

  int rg_parts() const override
  {
    if (kind == QMCB_C2C)
    {
      using SH = Shape<ST, true>;
      return ((2 * n_orb + SH::TILE - 1) / SH::TILE) * (SH::TILE / SH::VEC / 32);
    }
    using SH = Shape<ST, false>;
    return ((n_orb + SH::TILE - 1) / SH::TILE) * (SH::TILE / SH::VEC / 32);
  }

  template<bool C2C, int MODE>
  void launch(int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv, const int* ref_dev, void* phi_dev,
              void* rg_dev, cudaStream_t st)
  {
    using SH            = Shape<ST, C2C>;
    constexpr int TILE  = SH::TILE, STAGES = SH::STAGES, VEC = SH::VEC, MINB = SH::MINB;
    const int ncomp     = C2C ? 2 * n_orb : n_orb; // real components that matter
    const int ntiles    = (ncomp + TILE - 1) / TILE;
    SplineArgs<ST, ST> A;
    A.nw         = nw;
    A.r          = static_cast<const ST*>(r_dev);
    A.invrow     = static_cast<const ST*>(invrow_dev);
    A.ref        = ref_dev;
    A.ld_inv     = (long long)ld_inv;
    A.phi_vgl    = static_cast<ST*>(phi_dev);
    A.rg_partial = static_cast<ST*>(rg_dev);
    A.nparts     = ntiles * (TILE / VEC / 32);
    A.pdl_early  = (g_pdl_mode & 4) ? 1 : 0;
    static const int evict = [] {
      const char* e = std::getenv("QMCB_SPL_EVICT");
      return e ? std::atoi(e) : 1; // measured: device sweep 54.3 -> 51.2 us per move (the boundary kernel's rows stay in L2)
    }();
    A.l2_evict_first = evict;
    auto kern            = spline_gather_kernel<ST, ST, TILE, STAGES, VEC, MODE, C2C, MINB>;
    constexpr size_t smem = SplineSmem<ST, TILE, STAGES, VEC>::BYTES;
    ensure_dynamic_smem(kern, smem);
    const int nunits = nw * ntiles;
    if (nunits == 0)
      return;
    const int grid_x = std::min(nunits, MINB * sm_count());
    launch_kernel(kern, dim3(grid_x), dim3(TILE / VEC + 32), smem, st, (g_pdl_mode & 3) >= 1, tmap, dev, A, ntiles);
    QMCB_LAUNCH_CHECK();
  }

  void evaluate_dev(int mode, int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv, const int* ref_dev,
                    void* phi_dev, void* rg_dev, cudaStream_t st) override
  {
    if (kind == QMCB_C2C)
    {
      if (mode == MODE_V)
        launch<true, MODE_V>(nw, r_dev, invrow_dev, ld_inv, ref_dev, phi_dev, rg_dev, st);
      else
        launch<true, MODE_VGL>(nw, r_dev, invrow_dev, ld_inv, ref_dev, phi_dev, rg_dev, st);
    }
    else
    {
      if (mode == MODE_V)
        launch<false, MODE_V>(nw, r_dev, invrow_dev, ld_inv, ref_dev, phi_dev, rg_dev, st);
      else
        launch<false, MODE_VGL>(nw, r_dev, invrow_dev, ld_inv, ref_dev, phi_dev, rg_dev, st);
    }
  }
};

SplineSPOBase* make_spline(int precision, int kind, const int grid[3], int n_orb, int n_spl, size_t npad,
                           const void* coefs_host, const double G[9], const int halfG[3], const double* kcart)
{
  if (precision == QMCB_MIXED)
    return new SplineSPO<float>(precision, kind, grid, n_orb, n_spl, npad, coefs_host, G, halfG, kcart);
  if (precision == QMCB_FULL)
    return new SplineSPO<double>(precision, kind, grid, n_orb, n_spl, npad, coefs_host, G, halfG, kcart);
  throw std::runtime_error("unknown precision code");
}

} // namespace qmcb
