// qmcpack_b200/csrc/spline.cu -- host side of the spline SPOSet: table upload, launch configuration.
#include "internal.h"
#include "spline.cuh"
#include <cstring>
#include <mutex>

namespace qmcb
{
std::atomic<unsigned long long> g_launch_count{0};

namespace
{
int sm_count()
{
  static int n = 0;
  if (!n)
  {
    int dev = 0;
    QMCB_CUDA(cudaGetDevice(&dev));
    QMCB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// tile/stage shapes per (storage type, kind); shared memory per CTA = STAGES * 64 * TILE * sizeof(ST)
template<typename ST, bool C2C>
struct Shape;
template<>
struct Shape<float, false>
{
  static constexpr int TILE = 192, STAGES = 4, VEC = 1;
}; // 48 KB / stage
template<>
struct Shape<double, false>
{
  static constexpr int TILE = 128, STAGES = 3, VEC = 1;
}; // 64 KB / stage
template<>
struct Shape<float, true>
{
  static constexpr int TILE = 256, STAGES = 3, VEC = 2;
}; // 64 KB / stage
template<>
struct Shape<double, true>
{
  static constexpr int TILE = 128, STAGES = 3, VEC = 2;
}; // 64 KB / stage
} // namespace

template<typename ST>
struct SplineSPO : SplineSPOBase
{
  DevBuf<ST> coefs, kc, mkk;
  SplineDev<ST> dev;
  // cross-tile reduction scratch (grown on demand; one evaluation at a time per handle+stream is the contract,
  // crowds own their own scratch through scratch_for())
  struct Scratch
  {
    DevBuf<ST> partial;
    DevBuf<unsigned> ticket;
    int nw_cap = 0;
  };
  std::mutex mtx;
  std::vector<std::pair<cudaStream_t, std::unique_ptr<Scratch>>> scratch;

  size_t table_bytes() const override { return coefs.bytes(); }

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

  Scratch& scratch_for(cudaStream_t st, int nw, int ntiles, int nred)
  {
    std::lock_guard<std::mutex> lk(mtx);
    Scratch* s = nullptr;
    for (auto& e : scratch)
      if (e.first == st)
        s = e.second.get();
    if (!s)
    {
      scratch.emplace_back(st, std::make_unique<Scratch>());
      s = scratch.back().second.get();
    }
    if (s->nw_cap < nw)
    {
      QMCB_CUDA(cudaStreamSynchronize(st));
      s->partial.alloc((size_t)nw * ntiles * nred);
      s->ticket.alloc(nw);
      s->nw_cap = nw;
    }
    return *s;
  }

  template<bool C2C, int MODE>
  void launch(int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv, const int* ref_dev, void* phi_dev,
              void* rg_dev, cudaStream_t st)
  {
    using SH            = Shape<ST, C2C>;
    constexpr int TILE  = SH::TILE, STAGES = SH::STAGES, VEC = SH::VEC;
    const int ncomp     = C2C ? 2 * n_orb : n_orb; // real components that matter
    const int ntiles    = (ncomp + TILE - 1) / TILE;
    constexpr int nred  = C2C ? 8 : 4;
    SplineArgs<ST, ST> A;
    A.nw         = nw;
    A.r          = static_cast<const ST*>(r_dev);
    A.invrow     = static_cast<const ST*>(invrow_dev);
    A.ref        = ref_dev;
    A.ld_inv     = (long long)ld_inv;
    A.phi_vgl    = static_cast<ST*>(phi_dev);
    A.ratio_grad = static_cast<ST*>(rg_dev);
    A.partial    = nullptr;
    A.ticket     = nullptr;
    if (rg_dev && ntiles > 1)
    {
      Scratch& s = scratch_for(st, nw, ntiles, nred);
      A.partial  = s.partial.p;
      A.ticket   = s.ticket.p;
    }
    auto kern            = spline_gather_kernel<ST, ST, TILE, STAGES, VEC, MODE, C2C>;
    constexpr size_t smem = SplineSmem<ST, TILE, STAGES, VEC>::BYTES;
    static bool attr_set = false;
    if (!attr_set)
    {
      QMCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    const int nunits = nw * ntiles;
    if (nunits == 0)
      return;
    const int grid_x = std::min(nunits, sm_count());
    kern<<<grid_x, TILE / VEC + 32, smem, st>>>(dev, A, ntiles);
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
