// =====================================================================================
// oracle/qmc_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's batched per-electron-move hot path (QMCPACK, mounted
// read-only at /root/reference).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
// `--impl reference` legs of bench.py may load this; the product (qmcpack_b200/) never does.
//
// Every block cites the reference file:line it follows ("ref:" paths are relative to
// /root/reference/src).  Nothing here is copied: the arithmetic is restated in plain loops so
// that (a) it builds with g++ alone, (b) each floating-point association that matters for
// parity is visible.
//
// Pinning: tests/test_oracle_golden.py checks this file against the golden vectors of the
// reference's own unit tests (spline2/tests/test_multi_spline.cpp, QMCWaveFunctions/tests/
// test_DiracMatrix.cpp, test_bspline_jastrow values) and, when oracle/_ref was built from the
// reference sources (oracle/Makefile), against the reference's own compiled kernels
// (spline2::evaluate_*_impl, DelayedUpdate<T>, DiracMatrix<T>).
//
// Precision model (ref: config.h OHMMS_PRECISION / MIXED_PRECISION):
//   RT  = QMCTraits::RealType   positions, distance tables, Jastrows, driver scalars
//   ST  = spline storage type   (sposet precision="single|double")
//   VT  = QMCTraits::ValueType  orbital values, psiMinv, dpsiM, d2psiM
//   double = FullPrecRealType / PsiValue / matrix inversion (FPVT)
// full precision: RT=ST=VT=double;  mixed precision: RT=ST=VT=float.
// =====================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace orc
{
// -------------------------------------------------------------------------------------
// ref: Platforms/CPU/SIMD/aligned_allocator.hpp:41-47  getAlignedSize<T>(n): round n up to a
// multiple of QMC_SIMD_ALIGNMENT(64 B)/sizeof(T).
template<typename T>
inline size_t aligned_size(size_t n)
{
  constexpr size_t ND = 64 / sizeof(T);
  return ((n + ND - 1) / ND) * ND;
}

// 64-byte aligned std::vector (the reference kernels compiled into oracle/_ref use aligned SIMD access,
// `#pragma omp simd aligned(...: QMC_SIMD_ALIGNMENT)`, so every buffer handed to them must be aligned)
template<typename T>
struct AlignedAlloc
{
  using value_type = T;
  AlignedAlloc() = default;
  template<class U>
  AlignedAlloc(const AlignedAlloc<U>&)
  {}
  T* allocate(size_t n)
  {
    void* p = nullptr;
    if (posix_memalign(&p, 64, std::max<size_t>(64, ((n * sizeof(T) + 63) / 64) * 64)))
      throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { free(p); }
  template<class U>
  bool operator==(const AlignedAlloc<U>&) const
  {
    return true;
  }
  template<class U>
  bool operator!=(const AlignedAlloc<U>&) const
  {
    return false;
  }
};
template<typename T>
using avec = std::vector<T, AlignedAlloc<T>>;

// -------------------------------------------------------------------------------------
// ref: Numerics/SplineBound.hpp:37-62.  T is the type of the scaled coordinate (double for the
// 3-D tables because Ugrid::delta_inv is a double, MultiBsplineEval_helper.hpp:80-88), TR the
// residual type (ST).
template<typename T, typename TR>
inline void getSplineBound(const T x, const int nmax, int& ind, TR& dx)
{
  if (x < 0)
  {
    ind = 0;
    dx  = T(0);
  }
  else
  {
    T ipart;
    dx  = std::modf(x, &ipart);
    ind = static_cast<int>(ipart);
    if (ind > nmax)
    {
      ind = nmax;
      dx  = T(1) - std::numeric_limits<T>::epsilon();
    }
  }
}

// ref: spline2/MultiBsplineData.hpp:30-87 (A-matrix constants and Horner forms)
template<typename T>
inline void compute_prefactors(T a[4], T tx)
{
  constexpr T A00 = -1.0 / 6.0, A01 = 3.0 / 6.0, A02 = -3.0 / 6.0, A03 = 1.0 / 6.0;
  constexpr T A10 = 3.0 / 6.0, A11 = -6.0 / 6.0, A12 = 0.0 / 6.0, A13 = 4.0 / 6.0;
  constexpr T A20 = -3.0 / 6.0, A21 = 3.0 / 6.0, A22 = 3.0 / 6.0, A23 = 1.0 / 6.0;
  constexpr T A30 = 1.0 / 6.0, A31 = 0.0 / 6.0, A32 = 0.0 / 6.0, A33 = 0.0 / 6.0;
  a[0] = ((A00 * tx + A01) * tx + A02) * tx + A03;
  a[1] = ((A10 * tx + A11) * tx + A12) * tx + A13;
  a[2] = ((A20 * tx + A21) * tx + A22) * tx + A23;
  a[3] = ((A30 * tx + A31) * tx + A32) * tx + A33;
}

template<typename T>
inline void compute_prefactors(T a[4], T da[4], T d2a[4], T tx)
{
  compute_prefactors(a, tx);
  constexpr T dA01 = -0.5, dA02 = 1.0, dA03 = -0.5;
  constexpr T dA11 = 1.5, dA12 = -2.0, dA13 = 0.0;
  constexpr T dA21 = -1.5, dA22 = 1.0, dA23 = 0.5;
  constexpr T dA31 = 0.5, dA32 = 0.0, dA33 = 0.0;
  constexpr T d2A02 = -1.0, d2A03 = 1.0;
  constexpr T d2A12 = 3.0, d2A13 = -2.0;
  constexpr T d2A22 = -3.0, d2A23 = 1.0;
  constexpr T d2A32 = 1.0, d2A33 = 0.0;
  da[0]  = (dA01 * tx + dA02) * tx + dA03;
  da[1]  = (dA11 * tx + dA12) * tx + dA13;
  da[2]  = (dA21 * tx + dA22) * tx + dA23;
  da[3]  = (dA31 * tx + dA32) * tx + dA33;
  d2a[0] = d2A02 * tx + d2A03;
  d2a[1] = d2A12 * tx + d2A13;
  d2a[2] = d2A22 * tx + d2A23;
  d2a[3] = d2A32 * tx + d2A33;
}

// ref: QMCWaveFunctions/BsplineFactory/contraction_helper.hpp:47-48
template<typename T>
inline T SymTrace(T h00, T h01, T h02, T h11, T h12, T h22, const T gg[6])
{
  return h00 * gg[0] + h01 * gg[1] + h02 * gg[2] + h11 * gg[3] + h12 * gg[4] + h22 * gg[5];
}

// -------------------------------------------------------------------------------------
// Table descriptor: ref einspline/multi_bspline_structs.h (multi_UBspline_3d_{s,d}) with the
// periodic metadata of spline2/MultiBsplineBase.hpp:75-128: N_a = M_a + 3 planes, grid [0,1],
// delta = 1/M, strides x=Ny*Nz*Npad, y=Nz*Npad, z=Npad.  The table is NOT owned.
template<typename ST>
struct SplineTable
{
  const ST* coefs = nullptr;
  int M[3]        = {0, 0, 0}; // grid.num per axis
  int ns          = 0;         // num_splines (real components)
  size_t npad     = 0;         // z_stride
  size_t xs = 0, ys = 0, zs = 0;
  double start[3]     = {0, 0, 0};
  double delta_inv[3] = {0, 0, 0};

  void set(const ST* c, const int m[3], int nspl, size_t npad_)
  {
    coefs = c;
    ns    = nspl;
    npad  = npad_;
    for (int d = 0; d < 3; ++d)
    {
      M[d]               = m[d];
      const int Nd       = m[d] + 3;
      const double delta = (1.0 - 0.0) / (double)(Nd - 3);
      delta_inv[d]       = 1.0 / delta;
    }
    zs = npad;
    ys = (size_t)(M[2] + 3) * npad;
    xs = (size_t)(M[1] + 3) * (size_t)(M[2] + 3) * npad;
  }
  size_t coefs_size() const { return (size_t)(M[0] + 3) * xs; }
};

// ref: spline2/MultiBsplineEval_helper.hpp:32-93 computeLocationAndFractional
template<typename ST>
inline void locate(const SplineTable<ST>& s, ST x, ST y, ST z, int& ix, int& iy, int& iz, ST& tx, ST& ty, ST& tz)
{
  x -= s.start[0];
  y -= s.start[1];
  z -= s.start[2];
  getSplineBound(x * s.delta_inv[0], s.M[0] - 1, ix, tx);
  getSplineBound(y * s.delta_inv[1], s.M[1] - 1, iy, ty);
  getSplineBound(z * s.delta_inv[2], s.M[2] - 1, iz, tz);
}

// ref: spline2/MultiBsplineValue.hpp:21-54 evaluate_v_impl
template<typename ST>
inline void evaluate_v(const SplineTable<ST>& s, ST x, ST y, ST z, ST* vals)
{
  int ix, iy, iz;
  ST tx, ty, tz, a[4], b[4], c[4];
  locate(s, x, y, z, ix, iy, iz, tx, ty, tz);
  compute_prefactors(a, tx);
  compute_prefactors(b, ty);
  compute_prefactors(c, tz);
  const int n = s.ns;
  std::fill(vals, vals + n, ST());
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
    {
      const ST pre00 = a[i] * b[j];
      const ST* c0   = s.coefs + ((ix + i) * s.xs + (iy + j) * s.ys + iz * s.zs);
      const ST *c1 = c0 + s.zs, *c2 = c0 + 2 * s.zs, *c3 = c0 + 3 * s.zs;
      for (int m = 0; m < n; m++)
        vals[m] += pre00 * (c[0] * c0[m] + c[1] * c1[m] + c[2] * c2[m] + c[3] * c3[m]);
    }
}

// ref: spline2/MultiBsplineVGLH.hpp:29-117 evaluate_vgl_impl (grads [3][stride], lapl [3][stride];
// lapl[0] holds the lattice-unit Laplacian on exit)
template<typename ST>
inline void evaluate_vgl(const SplineTable<ST>& s, ST x, ST y, ST z, ST* vals, ST* grads, ST* lapl, size_t stride)
{
  int ix, iy, iz;
  ST tx, ty, tz, a[4], b[4], c[4], da[4], db[4], dc[4], d2a[4], d2b[4], d2c[4];
  locate(s, x, y, z, ix, iy, iz, tx, ty, tz);
  compute_prefactors(a, da, d2a, tx);
  compute_prefactors(b, db, d2b, ty);
  compute_prefactors(c, dc, d2c, tz);
  const int n = s.ns;
  ST *gx = grads, *gy = grads + stride, *gz = grads + 2 * stride;
  ST *lx = lapl, *ly = lapl + stride, *lz = lapl + 2 * stride;
  for (ST* p : {vals, gx, gy, gz, lx, ly, lz})
    std::fill(p, p + n, ST());
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
    {
      const ST pre20 = d2a[i] * b[j], pre10 = da[i] * b[j], pre00 = a[i] * b[j];
      const ST pre01 = a[i] * db[j], pre02 = a[i] * d2b[j];
      const ST* c0 = s.coefs + ((ix + i) * s.xs + (iy + j) * s.ys + iz * s.zs);
      const ST *c1 = c0 + s.zs, *c2 = c0 + 2 * s.zs, *c3 = c0 + 3 * s.zs;
      for (int m = 0; m < n; m++)
      {
        const ST v0 = c0[m], v1 = c1[m], v2 = c2[m], v3 = c3[m];
        const ST sum0 = c[0] * v0 + c[1] * v1 + c[2] * v2 + c[3] * v3;
        const ST sum1 = dc[0] * v0 + dc[1] * v1 + dc[2] * v2 + dc[3] * v3;
        const ST sum2 = d2c[0] * v0 + d2c[1] * v1 + d2c[2] * v2 + d2c[3] * v3;
        gx[m] += pre10 * sum0;
        gy[m] += pre01 * sum0;
        gz[m] += pre00 * sum1;
        lx[m] += pre20 * sum0;
        ly[m] += pre02 * sum0;
        lz[m] += pre00 * sum2;
        vals[m] += pre00 * sum0;
      }
    }
  const ST dxInv = s.delta_inv[0], dyInv = s.delta_inv[1], dzInv = s.delta_inv[2];
  const ST dxInv2 = dxInv * dxInv, dyInv2 = dyInv * dyInv, dzInv2 = dzInv * dzInv;
  for (int m = 0; m < n; m++)
  {
    gx[m] *= dxInv;
    gy[m] *= dyInv;
    gz[m] *= dzInv;
    lx[m] = lx[m] * dxInv2 + ly[m] * dyInv2 + lz[m] * dzInv2;
  }
}

// ref: spline2/MultiBsplineVGLH.hpp:120-227 evaluate_vgh_impl (grads [3][stride], hess [6][stride]
// in the order xx,xy,xz,yy,yz,zz); accumulation order i outer / j inner, prefactor products
// formed first (`pre = a[i]*b[j]`), z-sum written left to right.
template<typename ST>
inline void evaluate_vgh(const SplineTable<ST>& s, ST x, ST y, ST z, ST* vals, ST* grads, ST* hess, size_t stride)
{
  int ix, iy, iz;
  ST tx, ty, tz, a[4], b[4], c[4], da[4], db[4], dc[4], d2a[4], d2b[4], d2c[4];
  locate(s, x, y, z, ix, iy, iz, tx, ty, tz);
  compute_prefactors(a, da, d2a, tx);
  compute_prefactors(b, db, d2b, ty);
  compute_prefactors(c, dc, d2c, tz);
  const int n = s.ns;
  ST *gx = grads, *gy = grads + stride, *gz = grads + 2 * stride;
  ST *hxx = hess, *hxy = hess + stride, *hxz = hess + 2 * stride, *hyy = hess + 3 * stride, *hyz = hess + 4 * stride,
     *hzz = hess + 5 * stride;
  for (ST* p : {vals, gx, gy, gz, hxx, hxy, hxz, hyy, hyz, hzz})
    std::fill(p, p + n, ST());
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
    {
      const ST* c0 = s.coefs + ((ix + i) * s.xs + (iy + j) * s.ys + iz * s.zs);
      const ST *c1 = c0 + s.zs, *c2 = c0 + 2 * s.zs, *c3 = c0 + 3 * s.zs;
      const ST pre20 = d2a[i] * b[j], pre10 = da[i] * b[j], pre00 = a[i] * b[j];
      const ST pre11 = da[i] * db[j], pre01 = a[i] * db[j], pre02 = a[i] * d2b[j];
      for (int m = 0; m < n; m++)
      {
        const ST v0 = c0[m], v1 = c1[m], v2 = c2[m], v3 = c3[m];
        const ST sum0 = c[0] * v0 + c[1] * v1 + c[2] * v2 + c[3] * v3;
        const ST sum1 = dc[0] * v0 + dc[1] * v1 + dc[2] * v2 + dc[3] * v3;
        const ST sum2 = d2c[0] * v0 + d2c[1] * v1 + d2c[2] * v2 + d2c[3] * v3;
        hxx[m] += pre20 * sum0;
        hxy[m] += pre11 * sum0;
        hxz[m] += pre10 * sum1;
        hyy[m] += pre02 * sum0;
        hyz[m] += pre01 * sum1;
        hzz[m] += pre00 * sum2;
        gx[m] += pre10 * sum0;
        gy[m] += pre01 * sum0;
        gz[m] += pre00 * sum1;
        vals[m] += pre00 * sum0;
      }
    }
  const ST dxInv = s.delta_inv[0], dyInv = s.delta_inv[1], dzInv = s.delta_inv[2];
  const ST dxx = dxInv * dxInv, dyy = dyInv * dyInv, dzz = dzInv * dzInv;
  const ST dxy = dxInv * dyInv, dxz = dxInv * dzInv, dyz = dyInv * dzInv;
  for (int m = 0; m < n; m++)
  {
    gx[m] *= dxInv;
    gy[m] *= dyInv;
    gz[m] *= dzInv;
    hxx[m] *= dxx;
    hyy[m] *= dyy;
    hzz[m] *= dzz;
    hxy[m] *= dxy;
    hxz[m] *= dxz;
    hyz[m] *= dyz;
  }
}

// -------------------------------------------------------------------------------------
// Periodic interpolating coefficients.  ref: einspline/bspline_create.c:991-1035
// (solve_periodic_interp_1d_d), :1098-1120 (find_coefs_1d_d periodic branch, basis {1/6,2/3,1/6}),
// :1311-1390 (create_UBspline_3d_d: x, then y, then z sweeps).  data is [Mx][My][Mz] doubles,
// coefs is [Mx+3][My+3][Mz+3] doubles.
inline void solve_periodic_1d(const double* data, intptr_t dstride, double* coefs, intptr_t cstride, int M)
{
  std::vector<double> bands(4 * (size_t)M), lastCol(M);
  for (int i = 0; i < M; i++)
  {
    bands[4 * i + 0] = 1.0 / 6.0;
    bands[4 * i + 1] = 2.0 / 3.0;
    bands[4 * i + 2] = 1.0 / 6.0;
    bands[4 * i + 3] = data[i * dstride];
  }
  auto B = [&](int r, int c) -> double& { return bands[4 * r + c]; };
  B(0, 2) /= B(0, 1);
  B(0, 0) /= B(0, 1);
  B(0, 3) /= B(0, 1);
  B(0, 1) = 1.0;
  B(M - 1, 1) -= B(M - 1, 2) * B(0, 0);
  B(M - 1, 3) -= B(M - 1, 2) * B(0, 3);
  B(M - 1, 2) = -B(M - 1, 2) * B(0, 2);
  lastCol[0]  = B(0, 0);
  for (int row = 1; row < (M - 1); row++)
  {
    B(row, 1) -= B(row, 0) * B(row - 1, 2);
    B(row, 3) -= B(row, 0) * B(row - 1, 3);
    lastCol[row] = -B(row, 0) * lastCol[row - 1];
    B(row, 0)    = 0.0;
    B(row, 2) /= B(row, 1);
    B(row, 3) /= B(row, 1);
    lastCol[row] /= B(row, 1);
    B(row, 1) = 1.0;
    if (row < (M - 2))
    {
      B(M - 1, 3) -= B(M - 1, 2) * B(row, 3);
      B(M - 1, 1) -= B(M - 1, 2) * lastCol[row];
      B(M - 1, 2) = -B(M - 1, 2) * B(row, 2);
    }
  }
  B(M - 1, 0) += B(M - 1, 2);
  B(M - 1, 1) -= B(M - 1, 0) * (B(M - 2, 2) + lastCol[M - 2]);
  B(M - 1, 3) -= B(M - 1, 0) * B(M - 2, 3);
  B(M - 1, 3) /= B(M - 1, 1);
  coefs[M * cstride] = B(M - 1, 3);
  for (int row = M - 2; row >= 0; row--)
    coefs[(row + 1) * cstride] = B(row, 3) - B(row, 2) * coefs[(row + 2) * cstride] - lastCol[row] * coefs[M * cstride];
  coefs[0 * cstride]       = coefs[M * cstride];
  coefs[(M + 1) * cstride] = coefs[1 * cstride];
  coefs[(M + 2) * cstride] = coefs[2 * cstride];
}

inline void create_periodic_coefs_3d(const int M[3], const double* data, double* coefs)
{
  const int Mx = M[0], My = M[1], Mz = M[2];
  const int Nx = Mx + 3, Ny = My + 3, Nz = Mz + 3;
  for (int iy = 0; iy < My; iy++)
    for (int iz = 0; iz < Mz; iz++)
      solve_periodic_1d(data + (iy * Mz + iz), (intptr_t)My * Mz, coefs + (iy * Nz + iz), (intptr_t)Ny * Nz, Mx);
  // the in-place y and z sweeps read their data from coefs; copy the line first because the
  // periodic solve writes coefs[(row+1)*stride] while the band matrix already holds the data.
  for (int ix = 0; ix < Nx; ix++)
    for (int iz = 0; iz < Nz; iz++)
    {
      double* line = coefs + ((size_t)ix * Ny * Nz + iz);
      solve_periodic_1d(line, Nz, line, Nz, My);
    }
  for (int ix = 0; ix < Nx; ix++)
    for (int iy = 0; iy < Ny; iy++)
    {
      double* line = coefs + ((size_t)ix * Ny + iy) * Nz;
      solve_periodic_1d(line, 1, line, 1, Mz);
    }
}

// =====================================================================================
// SPO sets
// =====================================================================================
struct LatticeG
{
  double G[9];   // prim_lattice.G row-major, ru = r . G   (ref: Particle/Lattice/CrystalLattice.h:182-185)
  double GGt[9]; // dot(transpose(G), G)                    (ref: BsplineFactory/SplineR2R.cpp:40)
  int halfG[3] = {0, 0, 0};
  void set(const double g[9], const int hg[3])
  {
    for (int i = 0; i < 9; ++i)
      G[i] = g[i];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
      {
        double s = 0;
        for (int k = 0; k < 3; ++k)
          s += G[k * 3 + i] * G[k * 3 + j];
        GGt[i * 3 + j] = s;
      }
    for (int i = 0; i < 3; ++i)
      halfG[i] = hg ? hg[i] : 0;
  }
};

// ref: BsplineFactory/SplineR2R.h:156-170 convertPos; r is the Cartesian position in RT.
template<typename ST, typename RT>
inline int convertPos(const LatticeG& lat, const RT r[3], ST ru[3])
{
  // PointType is TinyVector<ST,3>: the position is first cast to ST (SplineR2R.cpp evaluateVGL: P.activeR(iat))
  // then dot(r, G) is evaluated with the ST-typed lattice.
  int bc_sign = 0;
  for (int j = 0; j < 3; ++j)
  {
    ST v = ST(0);
    for (int i = 0; i < 3; ++i)
      v += ST(r[i]) * ST(lat.G[i * 3 + j]);
    ru[j] = v;
  }
  for (int i = 0; i < 3; ++i)
    if (-std::numeric_limits<ST>::epsilon() < ru[i] && ru[i] < 0)
      ru[i] = ST(0.0);
    else
    {
      ST img = std::floor(ru[i]);
      ru[i] -= img;
      bc_sign += lat.halfG[i] * (int)img;
    }
  return bc_sign;
}

template<typename ST, typename VT>
struct SplineR2R
{
  SplineTable<ST> tab;
  LatticeG lat;
  int norb = 0; // OrbitalSetSize (<= tab.ns)
  avec<ST> myV, myG, myH;

  void init(const SplineTable<ST>& t, const LatticeG& l, int n)
  {
    tab  = t;
    lat  = l;
    norb = n;
    myV.assign(tab.npad, ST());
    myG.assign(3 * tab.npad, ST());
    myH.assign(6 * tab.npad, ST());
  }

  // ref: SplineR2R.cpp:100-118 evaluateValue + assign_v (:88-98)
  template<typename RT>
  void evaluateValue(const RT r[3], VT* psi)
  {
    ST ru[3];
    const int bc_sign = convertPos<ST, RT>(lat, r, ru);
    evaluate_v(tab, ru[0], ru[1], ru[2], myV.data());
    const ST signed_one = (bc_sign & 1) ? -1 : 1;
    for (int j = 0; j < norb; ++j)
      psi[j] = signed_one * myV[j];
  }

  // ref: SplineR2R.cpp:338-374 assign_vgl and :400-412 evaluateVGL.  dpsi is AoS [norb][3].
  template<typename RT>
  void evaluateVGL(const RT r[3], VT* psi, VT* dpsi, VT* d2psi)
  {
    ST ru[3];
    const int bc_sign = convertPos<ST, RT>(lat, r, ru);
    const size_t np   = tab.npad;
    evaluate_vgh(tab, ru[0], ru[1], ru[2], myV.data(), myG.data(), myH.data(), np);
    const ST signed_one = (bc_sign & 1) ? -1 : 1;
    const ST g00 = lat.G[0], g01 = lat.G[1], g02 = lat.G[2], g10 = lat.G[3], g11 = lat.G[4], g12 = lat.G[5],
             g20 = lat.G[6], g21 = lat.G[7], g22 = lat.G[8];
    const ST symGG[6] = {ST(lat.GGt[0]), ST(lat.GGt[1]) + ST(lat.GGt[3]), ST(lat.GGt[2]) + ST(lat.GGt[6]),
                         ST(lat.GGt[4]), ST(lat.GGt[5]) + ST(lat.GGt[7]), ST(lat.GGt[8])};
    const ST *g0 = myG.data(), *g1 = g0 + np, *g2 = g0 + 2 * np;
    const ST *h00 = myH.data(), *h01 = h00 + np, *h02 = h00 + 2 * np, *h11 = h00 + 3 * np, *h12 = h00 + 4 * np,
             *h22 = h00 + 5 * np;
    for (int j = 0; j < norb; ++j)
    {
      psi[j]          = signed_one * myV[j];
      dpsi[3 * j + 0] = signed_one * (g00 * g0[j] + g01 * g1[j] + g02 * g2[j]);
      dpsi[3 * j + 1] = signed_one * (g10 * g0[j] + g11 * g1[j] + g12 * g2[j]);
      dpsi[3 * j + 2] = signed_one * (g20 * g0[j] + g21 * g1[j] + g22 * g2[j]);
      d2psi[j]        = signed_one * SymTrace(h00[j], h01[j], h02[j], h11[j], h12[j], h22[j], symGG);
    }
  }
  int size() const { return norb; }
};

// ref: BsplineFactory/SplineC2C.cpp:200-277 assign_vgl (same arithmetic in ApplyPhaseC2C.hpp:48-118): complex orbital j
// from the real component pair (2j, 2j+1) of the VGH scratch, twist phase from the CARTESIAN position r
template<typename ST, typename VT>
inline void c2c_assign_vgl(const LatticeG& lat, size_t np, int norb, const ST* myV, const ST* myG, const ST* myH, const ST* kx,
                           const ST* ky, const ST* kz, const ST* mKK, const ST r[3], VT* psi, VT* dpsi, VT* d2psi)
{
  constexpr ST two(2);
  const ST g00 = lat.G[0], g01 = lat.G[1], g02 = lat.G[2], g10 = lat.G[3], g11 = lat.G[4], g12 = lat.G[5],
           g20 = lat.G[6], g21 = lat.G[7], g22 = lat.G[8];
  const ST x = r[0], y = r[1], z = r[2];
  const ST symGG[6] = {ST(lat.GGt[0]), ST(lat.GGt[1]) + ST(lat.GGt[3]), ST(lat.GGt[2]) + ST(lat.GGt[6]),
                       ST(lat.GGt[4]), ST(lat.GGt[5]) + ST(lat.GGt[7]), ST(lat.GGt[8])};
  const ST *g0 = myG, *g1 = g0 + np, *g2 = g0 + 2 * np;
  const ST *h00 = myH, *h01 = h00 + np, *h02 = h00 + 2 * np, *h11 = h00 + 3 * np, *h12 = h00 + 4 * np,
           *h22 = h00 + 5 * np;
  for (int j = 0; j < norb; ++j)
  {
    const size_t jr = 2 * (size_t)j, ji = jr + 1;
    const ST kX = kx[j], kY = ky[j], kZ = kz[j];
    const ST val_r = myV[jr], val_i = myV[ji];
    const ST ph = -(x * kX + y * kY + z * kZ);
    const ST s = std::sin(ph), c = std::cos(ph);
    const ST dX_r = g00 * g0[jr] + g01 * g1[jr] + g02 * g2[jr];
    const ST dY_r = g10 * g0[jr] + g11 * g1[jr] + g12 * g2[jr];
    const ST dZ_r = g20 * g0[jr] + g21 * g1[jr] + g22 * g2[jr];
    const ST dX_i = g00 * g0[ji] + g01 * g1[ji] + g02 * g2[ji];
    const ST dY_i = g10 * g0[ji] + g11 * g1[ji] + g12 * g2[ji];
    const ST dZ_i = g20 * g0[ji] + g21 * g1[ji] + g22 * g2[ji];
    const ST gX_r = dX_r + val_i * kX, gY_r = dY_r + val_i * kY, gZ_r = dZ_r + val_i * kZ;
    const ST gX_i = dX_i - val_r * kX, gY_i = dY_i - val_r * kY, gZ_i = dZ_i - val_r * kZ;
    const ST lcart_r = SymTrace(h00[jr], h01[jr], h02[jr], h11[jr], h12[jr], h22[jr], symGG);
    const ST lcart_i = SymTrace(h00[ji], h01[ji], h02[ji], h11[ji], h12[ji], h22[ji], symGG);
    const ST lap_r   = lcart_r + mKK[j] * val_r + two * (kX * dX_i + kY * dY_i + kZ * dZ_i);
    const ST lap_i   = lcart_i + mKK[j] * val_i - two * (kX * dX_r + kY * dY_r + kZ * dZ_r);
    psi[j]           = VT(c * val_r - s * val_i, c * val_i + s * val_r);
    dpsi[3 * j + 0]  = VT(c * gX_r - s * gX_i, c * gX_i + s * gX_r);
    dpsi[3 * j + 1]  = VT(c * gY_r - s * gY_i, c * gY_i + s * gY_r);
    dpsi[3 * j + 2]  = VT(c * gZ_r - s * gZ_i, c * gZ_i + s * gZ_r);
    d2psi[j]         = VT(c * lap_r - s * lap_i, c * lap_i + s * lap_r);
  }
}

// Complex orbitals from a table of 2*norb real components + twist phase.
// ref: BsplineFactory/SplineC2C.cpp:146-168 (assign_v), :200-277 (assign_vgl), identical math in
// ApplyPhaseC2C.hpp:20-118; kpoints: myKcart[j], mKK[j] = -|k_j|^2 (SplineC2C.h / BsplineSet.h:48-52).
template<typename ST, typename VT /* std::complex<..> */>
struct SplineC2C
{
  SplineTable<ST> tab;
  LatticeG lat;
  int norb = 0;
  std::vector<ST> kx, ky, kz, mKK;
  avec<ST> myV, myG, myH;

  void init(const SplineTable<ST>& t, const LatticeG& l, int n, const double* kcart /*[n][3]*/)
  {
    tab  = t;
    lat  = l;
    norb = n;
    kx.resize(n);
    ky.resize(n);
    kz.resize(n);
    mKK.resize(n);
    for (int j = 0; j < n; ++j)
    {
      kx[j]  = kcart[3 * j];
      ky[j]  = kcart[3 * j + 1];
      kz[j]  = kcart[3 * j + 2];
      mKK[j] = -(kcart[3 * j] * kcart[3 * j] + kcart[3 * j + 1] * kcart[3 * j + 1] + kcart[3 * j + 2] * kcart[3 * j + 2]);
    }
    myV.assign(tab.npad, ST());
    myG.assign(3 * tab.npad, ST());
    myH.assign(6 * tab.npad, ST());
  }

  // ref: CrystalLattice.h:187-198 toUnit_floor
  template<typename RT>
  void toUnit_floor(const RT r[3], ST ru[3]) const
  {
    for (int j = 0; j < 3; ++j)
    {
      ST v = ST(0);
      for (int i = 0; i < 3; ++i)
        v += ST(r[i]) * ST(lat.G[i * 3 + j]);
      ru[j] = v;
    }
    for (int i = 0; i < 3; i++)
      if (-std::numeric_limits<ST>::epsilon() < ru[i] && ru[i] < 0)
        ru[i] = ST(0.0);
      else
        ru[i] -= std::floor(ru[i]);
  }

  template<typename RT>
  void evaluateValue(const RT r[3], VT* psi)
  {
    ST ru[3];
    toUnit_floor(r, ru);
    evaluate_v(tab, ru[0], ru[1], ru[2], myV.data());
    const ST x = r[0], y = r[1], z = r[2];
    for (int j = 0; j < norb; ++j)
    {
      const ST val_r = myV[2 * j], val_i = myV[2 * j + 1];
      const ST ph = -(x * kx[j] + y * ky[j] + z * kz[j]);
      const ST s = std::sin(ph), c = std::cos(ph);
      psi[j] = VT(val_r * c - val_i * s, val_i * c + val_r * s);
    }
  }

  template<typename RT>
  void evaluateVGL(const RT r[3], VT* psi, VT* dpsi, VT* d2psi)
  {
    ST ru[3];
    toUnit_floor(r, ru);
    const size_t np = tab.npad;
    evaluate_vgh(tab, ru[0], ru[1], ru[2], myV.data(), myG.data(), myH.data(), np);
    const ST rc[3] = {ST(r[0]), ST(r[1]), ST(r[2])};
    c2c_assign_vgl<ST, VT>(lat, np, norb, myV.data(), myG.data(), myH.data(), kx.data(), ky.data(), kz.data(), mKK.data(), rc,
                           psi, dpsi, d2psi);
  }
  int size() const { return norb; }
};

// ref: QMCWaveFunctions/SPOSet.cpp:149-182 default mw_evaluateVGLandDetRatioGrads: per walker
// evaluateVGL, then ratio = dot(invRow, psi) (plain, not conjugated), grad = dot(invRow, dpsi)/ratio,
// and phi_vgl_v[5][nw][norb] filled field-major.
template<typename SPO, typename VT, typename RT>
inline void mw_evaluateVGLandDetRatioGrads(SPO& spo, int nw, const RT* r /*[nw][3]*/, const VT* const* invRow,
                                           VT* phi_vgl /*[5][nw][norb]*/, VT* ratios, VT* grads /*[nw][3]*/)
{
  const int n = spo.size();
  std::vector<VT> psi(n), dpsi(3 * (size_t)n), d2psi(n);
  const size_t stride = (size_t)nw * n;
  for (int iw = 0; iw < nw; ++iw)
  {
    spo.evaluateVGL(r + 3 * iw, psi.data(), dpsi.data(), d2psi.data());
    VT ratio(0), gx(0), gy(0), gz(0);
    // simd::dot accumulates left to right
    for (int j = 0; j < n; ++j)
      ratio += invRow[iw][j] * psi[j];
    for (int j = 0; j < n; ++j)
    {
      gx += invRow[iw][j] * dpsi[3 * j];
      gy += invRow[iw][j] * dpsi[3 * j + 1];
      gz += invRow[iw][j] * dpsi[3 * j + 2];
    }
    ratios[iw]        = ratio;
    grads[3 * iw + 0] = gx / ratio;
    grads[3 * iw + 1] = gy / ratio;
    grads[3 * iw + 2] = gz / ratio;
    VT* out = phi_vgl + (size_t)iw * n;
    for (int j = 0; j < n; ++j)
    {
      out[j]              = psi[j];
      out[j + stride]     = dpsi[3 * j];
      out[j + 2 * stride] = dpsi[3 * j + 1];
      out[j + 3 * stride] = dpsi[3 * j + 2];
      out[j + 4 * stride] = d2psi[j];
    }
  }
}

// =====================================================================================
// Dense inverse + log-determinant (FPVT = double always).
// ref: QMCWaveFunctions/Fermion/DiracMatrix.h:100-107 computeLogDet, :151-170 computeInvertAndLog,
// :189-221 invert_transpose.  The reference calls LAPACK getrf/getri (third-party, un-vendored;
// CMakeLists.txt:462-471); restated here as unblocked partial-pivoting LU (dgetf2 algorithm) followed
// by inversion from the factors (dgetri algorithm: inv(U), then solve X*L = inv(U), then undo the column
// interchanges).  LAPACK is column-major: invMat handed in row-major as (psiM)^T is the column-major psiM,
// so the result read row-major is (psiM^-1)^T, the "psiMinv" of the determinant classes.
template<typename T>
inline void lu_invert_colmajor(T* a, int n, int lda, std::complex<double>& logdet)
{
  std::vector<int> piv(n);
  // getf2: for each column j find pivot, swap rows, scale, rank-1 update
  for (int j = 0; j < n; ++j)
  {
    int p    = j;
    auto big = std::abs(a[j + (size_t)j * lda]);
    for (int i = j + 1; i < n; ++i)
    {
      auto v = std::abs(a[i + (size_t)j * lda]);
      if (v > big)
      {
        big = v;
        p   = i;
      }
    }
    piv[j] = p + 1; // 1-based like LAPACK
    if (a[p + (size_t)j * lda] == T(0))
      throw std::runtime_error("oracle LU: singular matrix");
    if (p != j)
      for (int c = 0; c < n; ++c)
        std::swap(a[j + (size_t)c * lda], a[p + (size_t)c * lda]);
    const T inv = T(1) / a[j + (size_t)j * lda];
    for (int i = j + 1; i < n; ++i)
      a[i + (size_t)j * lda] *= inv;
    for (int c = j + 1; c < n; ++c)
    {
      const T ajc = a[j + (size_t)c * lda];
      for (int i = j + 1; i < n; ++i)
        a[i + (size_t)c * lda] -= a[i + (size_t)j * lda] * ajc;
    }
  }
  // ref DiracMatrix.h:100-107: logdet = sum log(complex(pivot[i]==i+1 ? diag : -diag))
  logdet = std::complex<double>();
  for (int i = 0; i < n; ++i)
  {
    const T d = a[i + (size_t)i * lda];
    logdet += std::log(std::complex<double>((piv[i] == i + 1) ? d : -d));
  }
  // inv(U) in place (upper triangular, non-unit)
  for (int j = 0; j < n; ++j)
  {
    a[j + (size_t)j * lda] = T(1) / a[j + (size_t)j * lda];
    const T ajj            = -a[j + (size_t)j * lda];
    // column j above the diagonal <- (already inverted leading block) * (old column j) * ajj
    std::vector<T> col(j);
    for (int i = 0; i < j; ++i)
      col[i] = a[i + (size_t)j * lda];
    for (int i = 0; i < j; ++i)
    {
      T s = T(0);
      for (int k = i; k < j; ++k)
        s += a[i + (size_t)k * lda] * col[k];
      a[i + (size_t)j * lda] = s * ajj;
    }
  }
  // solve inv(A)*L = inv(U) for inv(A): columns from n-2 down to 0
  std::vector<T> work(n);
  for (int j = n - 2; j >= 0; --j)
  {
    for (int i = j + 1; i < n; ++i)
    {
      work[i]                = a[i + (size_t)j * lda];
      a[i + (size_t)j * lda] = T(0);
    }
    for (int k = j + 1; k < n; ++k)
    {
      const T w = work[k];
      for (int i = 0; i < n; ++i)
        a[i + (size_t)j * lda] -= a[i + (size_t)k * lda] * w;
    }
  }
  // apply column interchanges in reverse
  for (int j = n - 2; j >= 0; --j)
  {
    const int jp = piv[j] - 1;
    if (jp != j)
      for (int i = 0; i < n; ++i)
        std::swap(a[i + (size_t)j * lda], a[i + (size_t)jp * lda]);
  }
}

// invert_transpose: amat is psiM [n][n_cols>=n] row-major (row = electron, col = orbital), invMat is
// [n][lda] row-major and receives (psiM^-1)^T, computed in double, cast to VT.
// FPVT of a value type: double for real VT, std::complex<double> for complex VT (Configuration.h FullPrecValueType)
template<typename VT>
struct FullPrec
{
  using type = double;
};
template<typename R>
struct FullPrec<std::complex<R>>
{
  using type = std::complex<double>;
};

template<typename VT>
inline void invert_transpose(const VT* amat, int n, int a_cols, VT* invMat, int lda, std::complex<double>& logdet)
{
  using FP = typename FullPrec<VT>::type;
  // simd::transpose(amat) -> psiM_fp[n][lda] row-major; that buffer read column-major by LAPACK is amat.
  std::vector<FP> fp((size_t)n * lda, FP(0));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      fp[(size_t)j * lda + i] = static_cast<FP>(amat[(size_t)i * a_cols + j]);
  lu_invert_colmajor(fp.data(), n, lda, logdet);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      invMat[(size_t)i * lda + j] = static_cast<VT>(fp[(size_t)i * lda + j]);
}

// =====================================================================================
// Delayed (rank-k) update engine.
// ref: QMCWaveFunctions/Fermion/DelayedUpdate.h:85-103 getInvRow, :113-141 acceptRow, :145-217
// updateInvMat.  BLAS is column-major there; everything below is written with explicit row-major
// indices:  Ainv[r][j] (row r = electron), U[c][j], V[c][j], Binv[a][b] = Binv_mem[a*k + b].
//   getInvRow:  p_c   = sum_j U[c][j] x_j                       (gemv 'T' on U)
//               w_b   = - sum_a Binv[a][b] p_a                  (gemv 'N' on Binv; stored in row `count` of Binv)
//               x_j  += sum_c V[c][j] w_c                       (gemv 'N' on V)
//   acceptRow:  p_a   = - sum_j V[a][j] psiV_j ; sigma = 1/ratio
//               y_a   = sigma * sum_b Binv[a][b] p_b -> Binv[a][count]   (gemv 'T' on Binv, output stride k)
//               Binv[a][b] += y_a * w_b   (a,b < count; w = row `count` left by getInvRow)   (ger)
//               Binv[count][b] = sigma * w_b ; Binv[count][count] = sigma
//   updateInvMat (count>1): T[m][c] = sum_j U[c][j] Ainv[m][j]  (tempMat[m][c]); T[list[c]][c] -= 1;
//               U'[c][j] = sum_a Binv[c][a] V[a][j];  Ainv[m][j] -= sum_c T[m][c] U'[c][j].
// The batched engine (DelayedUpdateBatched.h:542-670 + detail/OMPTarget/AccelMatrixUpdateOMPTarget.hpp:
// 96-163) adds the pseudo-accept for rejected walkers; see pseudoAcceptRow.
template<typename T>
struct DelayedUpdate
{
  int n = 0, k = 0;
  std::vector<T> U, V, Binv, tempMat, temp, p;
  std::vector<int> delay_list;
  int delay_count = 0;

  void resize(int norb, int delay)
  {
    n = norb;
    k = delay;
    V.assign((size_t)k * n, T());
    U.assign((size_t)k * n, T());
    p.assign(k, T());
    temp.assign(n, T());
    tempMat.assign((size_t)n * k, T());
    Binv.assign((size_t)k * k, T());
    delay_list.assign(k, 0);
    delay_count = 0;
  }

  void getInvRow(const T* Ainv, int lda, int rowchanged, T* invRow)
  {
    const T* arow = Ainv + (size_t)rowchanged * lda;
    std::copy_n(arow, n, invRow);
    if (delay_count == 0)
      return;
    const int c = delay_count;
    for (int a = 0; a < c; ++a)
    {
      T s(0);
      for (int j = 0; j < n; ++j)
        s += U[(size_t)a * n + j] * invRow[j];
      p[a] = s;
    }
    T* w = &Binv[(size_t)c * k];
    for (int b = 0; b < c; ++b)
    {
      T s(0);
      for (int a = 0; a < c; ++a)
        s += Binv[(size_t)a * k + b] * p[a];
      w[b] = -s;
    }
    for (int a = 0; a < c; ++a)
      for (int j = 0; j < n; ++j)
        invRow[j] += V[(size_t)a * n + j] * w[a];
  }

  template<typename RATIOT>
  void acceptRow(T* Ainv, int lda, int rowchanged, const T* psiV, const RATIOT ratio_new)
  {
    const int c = delay_count;
    std::copy_n(Ainv + (size_t)rowchanged * lda, n, &V[(size_t)c * n]);
    std::copy_n(psiV, n, &U[(size_t)c * n]);
    delay_list[c] = rowchanged;
    for (int a = 0; a <= c; ++a)
    {
      T s(0);
      for (int j = 0; j < n; ++j)
        s += V[(size_t)a * n + j] * psiV[j];
      p[a] = -s;
    }
    const T sigma            = static_cast<T>(RATIOT(1) / ratio_new);
    Binv[(size_t)c * k + c] = sigma;
    for (int a = 0; a < c; ++a)
    {
      T s(0);
      for (int b = 0; b < c; ++b)
        s += Binv[(size_t)a * k + b] * p[b];
      Binv[(size_t)a * k + c] = sigma * s;
    }
    for (int a = 0; a < c; ++a)
      for (int b = 0; b < c; ++b)
        Binv[(size_t)a * k + b] += Binv[(size_t)a * k + c] * Binv[(size_t)c * k + b];
    for (int b = 0; b < c; ++b)
      Binv[(size_t)c * k + b] *= sigma;
    delay_count++;
    if (delay_count == k)
      updateInvMat(Ainv, lda);
  }

  // ref: detail/OMPTarget/AccelMatrixUpdateOMPTarget.hpp:139-160 (rejected walker keeps the crowd's
  // delay_count in step): V[c] = Ainv[row] (DelayedUpdateBatched.h:646 copy for all walkers), U[c] = 0,
  // Binv row/col c = 0, diag = 1, delay_list[c] = -1.
  void pseudoAcceptRow(T* Ainv, int lda, int rowchanged)
  {
    const int c = delay_count;
    std::copy_n(Ainv + (size_t)rowchanged * lda, n, &V[(size_t)c * n]);
    std::fill_n(&U[(size_t)c * n], n, T(0));
    for (int b = 0; b < c; ++b)
      Binv[(size_t)c * k + b] = Binv[(size_t)b * k + c] = T(0);
    Binv[(size_t)c * k + c] = T(1);
    delay_list[c]           = -1;
    delay_count++;
    if (delay_count == k)
      updateInvMat(Ainv, lda);
  }

  void updateInvMat(T* Ainv, int lda)
  {
    if (delay_count == 0)
      return;
    const int c = delay_count;
    if (c == 1 && delay_list[0] >= 0)
    {
      // Fahy's variant of Sherman-Morrison, DelayedUpdate.h:156-162
      for (int m = 0; m < n; ++m)
      {
        T s(0);
        for (int j = 0; j < n; ++j)
          s += Ainv[(size_t)m * lda + j] * U[j];
        temp[m] = s;
      }
      temp[delay_list[0]] -= T(1);
      const T alpha = -Binv[0];
      for (int m = 0; m < n; ++m)
        for (int j = 0; j < n; ++j)
          Ainv[(size_t)m * lda + j] += alpha * V[j] * temp[m];
    }
    else
    {
      for (int m = 0; m < n; ++m)
        for (int a = 0; a < c; ++a)
        {
          T s(0);
          for (int j = 0; j < n; ++j)
            s += U[(size_t)a * n + j] * Ainv[(size_t)m * lda + j];
          tempMat[(size_t)m * k + a] = s;
        }
      for (int a = 0; a < c; ++a)
        if (delay_list[a] >= 0) // applyW_batched skips -1 (AccelMatrixUpdateOMPTarget.hpp:175-188)
          tempMat[(size_t)delay_list[a] * k + a] -= T(1);
      // U' overwrites U
      std::vector<T> Up((size_t)c * n);
      for (int a = 0; a < c; ++a)
        for (int j = 0; j < n; ++j)
        {
          T s(0);
          for (int b = 0; b < c; ++b)
            s += V[(size_t)b * n + j] * Binv[(size_t)a * k + b];
          Up[(size_t)a * n + j] = s;
        }
      std::copy(Up.begin(), Up.end(), U.begin());
      for (int m = 0; m < n; ++m)
        for (int j = 0; j < n; ++j)
        {
          T s(0);
          for (int a = 0; a < c; ++a)
            s += U[(size_t)a * n + j] * tempMat[(size_t)m * k + a];
          Ainv[(size_t)m * lda + j] -= s;
        }
    }
    delay_count = 0;
  }
};

// =====================================================================================
// Minimum-image distances.
// ref: Particle/Lattice/ParticleBConds3DSoa.h:111-139 (PPPO orthorhombic) and :390-449 (PPPG general
// cell; reduce with floor in the reduced basis then search the 8 cell corners).
template<typename T>
struct MinImage
{
  bool ortho = true;
  T L[3], Linv[3];
  T r[9], g[9];     // general cell: R rows, G = inverse(R)
  T corners[3][8];  // corners[idim][c]

  // ref: Particle/Lattice/LatticeAnalyzer.h:213-246 found_shorter_base
  static bool found_shorter_base(T rb[3][3])
  {
    const T eps = T(10) * std::numeric_limits<T>::epsilon();
    auto dot3   = [](const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    int imax    = 0;
    T r2max     = dot3(rb[0], rb[0]);
    for (int i = 1; i < 3; i++)
    {
      T r2 = dot3(rb[i], rb[i]);
      if ((r2 - r2max) > eps)
      {
        r2max = r2;
        imax  = i;
      }
    }
    T rmax = std::sqrt(r2max);
    T tol  = T(4.0) * rmax * eps;
    T rb_new[4][3];
    for (int d = 0; d < 3; ++d)
    {
      rb_new[0][d] = rb[0][d] + rb[1][d] - rb[2][d];
      rb_new[1][d] = rb[0][d] + rb[2][d] - rb[1][d];
      rb_new[2][d] = rb[1][d] + rb[2][d] - rb[0][d];
      rb_new[3][d] = rb[0][d] + rb[1][d] + rb[2][d];
    }
    for (int i = 0; i < 4; ++i)
    {
      T r2 = dot3(rb_new[i], rb_new[i]);
      if ((r2 - r2max) < -tol)
      {
        for (int d = 0; d < 3; ++d)
          rb[imax][d] = rb_new[i][d];
        return true;
      }
    }
    return false;
  }
  // ref: LatticeAnalyzer.h:247-272 find_reduced_basis
  static void find_reduced_basis(T rb[3][3])
  {
    const int maxIter = 10000;
    for (int count = 0; count < maxIter; count++)
    {
      T saved[3][3];
      for (int i = 0; i < 3; ++i)
        for (int d = 0; d < 3; ++d)
          saved[i][d] = rb[i][d];
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
    throw std::runtime_error("Reduced basis not found in allowed number of iterations.");
  }

  // ref: ParticleBConds3DSoa.h:339-386 (DTD_BConds<T,3,PPPG+SOA_OFFSET> constructor: reduced basis rb in T, g = inverse(rb)
  // in T, corners from rb) and :111-128 (PPPO: box lengths)
  void set(const double R[9])
  {
    ortho = (R[1] == 0 && R[2] == 0 && R[3] == 0 && R[5] == 0 && R[6] == 0 && R[7] == 0);
    T rb[3][3];
    for (int i = 0; i < 3; ++i)
      for (int d = 0; d < 3; ++d)
        rb[i][d] = (T)R[3 * i + d];
    if (!ortho)
      find_reduced_basis(rb);
    const T* a = &rb[0][0];
    // ref: OhmmsPETE/TensorOps.h:906-923 inverse(Tensor<T,3>)
    const T det  = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    const T vinv = T(1) / det;
    g[0]         = vinv * (a[4] * a[8] - a[5] * a[7]);
    g[1]         = vinv * (a[7] * a[2] - a[8] * a[1]);
    g[2]         = vinv * (a[1] * a[5] - a[2] * a[4]);
    g[3]         = vinv * (a[5] * a[6] - a[3] * a[8]);
    g[4]         = vinv * (a[8] * a[0] - a[6] * a[2]);
    g[5]         = vinv * (a[2] * a[3] - a[0] * a[5]);
    g[6]         = vinv * (a[3] * a[7] - a[4] * a[6]);
    g[7]         = vinv * (a[6] * a[1] - a[7] * a[0]);
    g[8]         = vinv * (a[0] * a[4] - a[1] * a[3]);
    for (int i = 0; i < 9; ++i)
      r[i] = a[i];
    for (int d = 0; d < 3; ++d)
    {
      const double len = std::sqrt(R[3 * d] * R[3 * d] + R[3 * d + 1] * R[3 * d + 1] + R[3 * d + 2] * R[3 * d + 2]);
      L[d]             = (T)len;
      Linv[d]          = (T)(1.0 / len);
    }
    // ref :371-385 corners: 0, -a0, -a1, -a2, -(a0+a1), -(a0+a2), -(a1+a2), -(a0+a1+a2) of the reduced basis
    for (int d = 0; d < 3; ++d)
    {
      const T a0 = rb[0][d], a1 = rb[1][d], a2 = rb[2][d];
      corners[d][0] = T(0);
      corners[d][1] = T(-1) * a0;
      corners[d][2] = T(-1) * a1;
      corners[d][3] = T(-1) * a2;
      corners[d][4] = T(-1) * (a0 + a1);
      corners[d][5] = T(-1) * (a0 + a2);
      corners[d][6] = T(-1) * (a1 + a2);
      corners[d][7] = T(-1) * (a0 + a1 + a2);
    }
  }

  // displacement source(px,py,pz) - pos, one pair
  inline void pair(const T pos[3], T px, T py, T pz, int iel, int flip_ind, T& rr, T& dx, T& dy, T& dz) const
  {
    if (ortho)
    {
      const T x = (px - pos[0]) * Linv[0];
      const T y = (py - pos[1]) * Linv[1];
      const T z = (pz - pos[2]) * Linv[2];
      dx        = L[0] * (x - std::round(x));
      dy        = L[1] * (y - std::round(y));
      dz        = L[2] * (z - std::round(z));
      rr        = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    else
    {
      const T flip    = iel < flip_ind ? T(1) : T(-1);
      const T displ_0 = (px - pos[0]) * flip;
      const T displ_1 = (py - pos[1]) * flip;
      const T displ_2 = (pz - pos[2]) * flip;
      const T ar_0    = -std::floor(displ_0 * g[0] + displ_1 * g[3] + displ_2 * g[6]);
      const T ar_1    = -std::floor(displ_0 * g[1] + displ_1 * g[4] + displ_2 * g[7]);
      const T ar_2    = -std::floor(displ_0 * g[2] + displ_1 * g[5] + displ_2 * g[8]);
      const T delx    = displ_0 + ar_0 * r[0] + ar_1 * r[3] + ar_2 * r[6];
      const T dely    = displ_1 + ar_0 * r[1] + ar_1 * r[4] + ar_2 * r[7];
      const T delz    = displ_2 + ar_0 * r[2] + ar_1 * r[5] + ar_2 * r[8];
      T rmin          = delx * delx + dely * dely + delz * delz;
      int ic          = 0;
      for (int c = 1; c < 8; ++c)
      {
        const T x  = delx + corners[0][c];
        const T y  = dely + corners[1][c];
        const T z  = delz + corners[2][c];
        const T r2 = x * x + y * y + z * z;
        ic         = (r2 < rmin) ? c : ic;
        rmin       = (r2 < rmin) ? r2 : rmin;
      }
      rr = std::sqrt(rmin);
      dx = flip * (delx + corners[0][ic]);
      dy = flip * (dely + corners[1][ic]);
      dz = flip * (delz + corners[2][ic]);
    }
  }

  // ref: SoaDistanceTableAAOMPTarget.h:306-364 rows; rsoa is [3][npad]; out is [4][npad] = r,dx,dy,dz
  void row(const T pos[3], const T* rsoa, size_t npad, int nsrc, int flip_ind, T* out) const
  {
    for (int j = 0; j < nsrc; ++j)
      pair(pos, rsoa[j], rsoa[npad + j], rsoa[2 * npad + j], j, flip_ind, out[j], out[npad + j], out[2 * npad + j],
           out[3 * npad + j]);
  }
};

// =====================================================================================
// 1-D cubic B-spline Jastrow functor.
// ref: QMCWaveFunctions/Jastrow/BsplineFunctor.h:102-131 (resize/reset: coefficient vector from the
// parameters and the cusp), :226-285 (evaluate_impl), A-matrix constants :47-66.
template<typename RT>
struct BsplineFunctor
{
  std::vector<RT> coefs;
  RT DeltaR = 0, DeltaRInv = 0, cutoff_radius = 0, CuspValue = 0;
  int max_index = 0;
  bool present  = false;

  void set(const double* params, int nparams, double rcut, double cusp)
  {
    present             = true;
    cutoff_radius       = (RT)rcut;
    CuspValue           = (RT)cusp;
    const int numCoefs  = nparams + 4;
    const int numKnots  = numCoefs - 2;
    DeltaR              = cutoff_radius / (RT)(numKnots - 1);
    DeltaRInv           = 1.0 / DeltaR;
    coefs.assign(numCoefs, RT(0));
    std::vector<RT> P(params, params + nparams);
    coefs[1] = P[0];
    coefs[2] = P[1];
    coefs[0] = P[1] - 2.0 * DeltaR * CuspValue;
    for (int i = 2; i < nparams; i++)
      coefs[i + 1] = P[i];
    max_index = numCoefs - 4;
  }

  inline RT evaluate_impl(RT r, RT& dudr, RT& d2udr2) const
  {
    constexpr RT A0 = -1.0 / 6.0, A1 = 3.0 / 6.0, A2 = -3.0 / 6.0, A3 = 1.0 / 6.0;
    constexpr RT A4 = 3.0 / 6.0, A5 = -6.0 / 6.0, A6 = 0.0 / 6.0, A7 = 4.0 / 6.0;
    constexpr RT A8 = -3.0 / 6.0, A9 = 3.0 / 6.0, A10 = 3.0 / 6.0, A11 = 1.0 / 6.0;
    constexpr RT A12 = 1.0 / 6.0, A13 = 0.0 / 6.0, A14 = 0.0 / 6.0, A15 = 0.0 / 6.0;
    constexpr RT dA1 = -0.5, dA2 = 1.0, dA3 = -0.5;
    constexpr RT dA5 = 1.5, dA6 = -2.0, dA7 = 0.0;
    constexpr RT dA9 = -1.5, dA10 = 1.0, dA11 = 0.5;
    constexpr RT dA13 = 0.5, dA14 = 0.0, dA15 = 0.0;
    constexpr RT d2A2 = -1.0, d2A3 = 1.0, d2A6 = 3.0, d2A7 = -2.0, d2A10 = -3.0, d2A11 = 1.0, d2A14 = 1.0, d2A15 = 0.0;
    r *= DeltaRInv;
    int i;
    RT t;
    getSplineBound(r, max_index, i, t);
    const RT sCoef0 = coefs[i + 0], sCoef1 = coefs[i + 1], sCoef2 = coefs[i + 2], sCoef3 = coefs[i + 3];
    d2udr2 = DeltaRInv * DeltaRInv *
        (sCoef0 * (d2A2 * t + d2A3) + sCoef1 * (d2A6 * t + d2A7) + sCoef2 * (d2A10 * t + d2A11) +
         sCoef3 * (d2A14 * t + d2A15));
    dudr = DeltaRInv *
        (sCoef0 * ((dA1 * t + dA2) * t + dA3) + sCoef1 * ((dA5 * t + dA6) * t + dA7) +
         sCoef2 * ((dA9 * t + dA10) * t + dA11) + sCoef3 * ((dA13 * t + dA14) * t + dA15));
    return (sCoef0 * (((A0 * t + A1) * t + A2) * t + A3) + sCoef1 * (((A4 * t + A5) * t + A6) * t + A7) +
            sCoef2 * (((A8 * t + A9) * t + A10) * t + A11) + sCoef3 * (((A12 * t + A13) * t + A14) * t + A15));
  }

  // ref BsplineFunctor.h:254-264 evaluate(r, dudr, d2udr2): zero beyond the cutoff
  inline RT evaluate(RT r, RT& dudr, RT& d2udr2) const
  {
    RT u(0);
    dudr   = RT(0);
    d2udr2 = RT(0);
    if (present && r < cutoff_radius)
      u = evaluate_impl(r, dudr, d2udr2);
    return u;
  }
  inline RT evaluate(RT r) const
  {
    RT a, b;
    return evaluate(r, a, b);
  }
};

// =====================================================================================
// Two-body Jastrow, batched (offload-style) form.
// ref: Jastrow/BsplineFunctor.cpp:26-131 mw_evaluateVGL, :203-326 mw_updateVGL;
// Jastrow/TwoBodyJastrow.cpp:542-578 mw_ratioGrad, :631-664 mw_accept_rejectMove, :667-713 recompute.
// The functor for pair (iat, j) is F[group(iat)*NumGroups + group(j)] (TwoBodyJastrow.cpp:433-443).
template<typename RT>
struct TwoBodyJastrow
{
  int N = 0, ngroups = 0;
  size_t npad = 0;
  std::vector<int> grp_ids;              // [N]
  std::vector<BsplineFunctor<RT>> F;     // [ngroups*ngroups]
  // per walker state
  struct State
  {
    std::vector<RT> Uat, dUat /*[3][npad]*/, d2Uat;
    std::vector<RT> cur_allu; // [3][npad] u, du/r, d2u of the proposed row
    RT cur_vgl[5];
    double log_value = 0;
  };

  void init(int n, int ng, const int* gids)
  {
    N       = n;
    ngroups = ng;
    npad    = aligned_size<RT>(n);
    grp_ids.assign(gids, gids + n);
    F.assign((size_t)ng * ng, BsplineFunctor<RT>());
  }
  void initState(State& s) const
  {
    s.Uat.assign(npad, RT(0));
    s.dUat.assign(3 * npad, RT(0));
    s.d2Uat.assign(npad, RT(0));
    s.cur_allu.assign(3 * npad, RT(0));
    s.log_value = 0;
  }

  // mw_evaluateVGL body for one walker; dist = temp row [4][npad]
  void evaluateVGL(State& s, int iat, const RT* dist) const
  {
    RT val_sum(0), grad_x(0), grad_y(0), grad_z(0), lapl(0);
    const RT *dx = dist + npad, *dy = dist + 2 * npad, *dz = dist + 3 * npad;
    const int igt = grp_ids[iat] * ngroups;
    for (int j = 0; j < N; ++j)
    {
      if (j == iat)
        continue;
      const BsplineFunctor<RT>& f = F[igt + grp_ids[j]];
      const RT r                  = dist[j];
      RT u(0), dudr(0), d2udr2(0);
      if (f.present && r < f.cutoff_radius)
      {
        u = f.evaluate_impl(r, dudr, d2udr2);
        dudr *= RT(1) / r;
      }
      s.cur_allu[j]            = u;
      s.cur_allu[j + npad]     = dudr;
      s.cur_allu[j + 2 * npad] = d2udr2;
      val_sum += u;
      lapl += d2udr2 + RT(2) * dudr;
      grad_x += dudr * dx[j];
      grad_y += dudr * dy[j];
      grad_z += dudr * dz[j];
    }
    s.cur_vgl[0] = val_sum;
    s.cur_vgl[1] = grad_x;
    s.cur_vgl[2] = grad_y;
    s.cur_vgl[3] = grad_z;
    s.cur_vgl[4] = -lapl;
  }

  // ratio (PsiValue=double) = exp(Uat[iat] - cur_Uat), grad_new += cur grad  (TwoBodyJastrow.cpp:570-576)
  double ratioGrad(State& s, int iat, const RT* dist_new, RT grad_new[3]) const
  {
    evaluateVGL(s, iat, dist_new);
    for (int d = 0; d < 3; ++d)
      grad_new[d] += s.cur_vgl[1 + d];
    return std::exp(static_cast<double>(s.Uat[iat] - s.cur_vgl[0]));
  }

  // accepted walker: mw_accept_rejectMove log update (:656-659) + mw_updateVGL body (BsplineFunctor.cpp:262-324)
  void acceptMove(State& s, int iat, const RT* dist_new, const RT* dist_old) const
  {
    s.log_value += s.Uat[iat] - s.cur_vgl[0];
    const RT *nx = dist_new + npad, *ny = dist_new + 2 * npad, *nz = dist_new + 3 * npad;
    const RT *ox = dist_old + npad, *oy = dist_old + 2 * npad, *oz = dist_old + 3 * npad;
    RT *dUx = s.dUat.data(), *dUy = dUx + npad, *dUz = dUx + 2 * npad;
    const int igt = grp_ids[iat] * ngroups;
    for (int j = 0; j < N; ++j)
    {
      if (j == iat)
        continue;
      const BsplineFunctor<RT>& f = F[igt + grp_ids[j]];
      const RT r                  = dist_old[j];
      RT u(0), dudr(0), d2udr2(0);
      if (f.present && r < f.cutoff_radius)
      {
        u = f.evaluate_impl(r, dudr, d2udr2);
        dudr *= RT(1) / r;
      }
      const RT cur_u = s.cur_allu[j], cur_dudr = s.cur_allu[j + npad], cur_d2udr2 = s.cur_allu[j + 2 * npad];
      s.Uat[j] += cur_u - u;
      dUx[j] -= nx[j] * cur_dudr - ox[j] * dudr;
      dUy[j] -= ny[j] * cur_dudr - oy[j] * dudr;
      dUz[j] -= nz[j] * cur_dudr - oz[j] * dudr;
      constexpr RT lapfac(2);
      s.d2Uat[j] -= cur_d2udr2 + lapfac * cur_dudr - (d2udr2 + lapfac * dudr);
    }
    s.Uat[iat]   = s.cur_vgl[0];
    dUx[iat]     = s.cur_vgl[1];
    dUy[iat]     = s.cur_vgl[2];
    dUz[iat]     = s.cur_vgl[3];
    s.d2Uat[iat] = s.cur_vgl[4];
  }

  // ref TwoBodyJastrow.cpp:667-713 recompute (lower-triangle accumulation) + :769-808 computeGL/log:
  // log_value = -0.5 * sum_i Uat[i].  rsoa [3][npad_pos].
  void recompute(State& s, const MinImage<RT>& mi, const RT* rsoa, size_t npad_pos) const
  {
    std::fill(s.Uat.begin(), s.Uat.end(), RT(0));
    std::fill(s.dUat.begin(), s.dUat.end(), RT(0));
    std::fill(s.d2Uat.begin(), s.d2Uat.end(), RT(0));
    std::vector<RT> row(4 * npad), u(N), du(N), d2u(N);
    RT *dUx = s.dUat.data(), *dUy = dUx + npad, *dUz = dUx + 2 * npad;
    for (int iat = 0; iat < N; ++iat)
    {
      // full-table row iat: displacement convention of SoaDistanceTableAA::evaluate (:108-115):
      // computeDistances(R[iat], all, ..., 0, iat, iat) -> d = R_j - R_iat for j < iat
      RT pos[3] = {rsoa[iat], rsoa[npad_pos + iat], rsoa[2 * npad_pos + iat]};
      for (int j = 0; j < iat; ++j)
        mi.pair(pos, rsoa[j], rsoa[npad_pos + j], rsoa[2 * npad_pos + j], j, iat, row[j], row[npad + j],
                row[2 * npad + j], row[3 * npad + j]);
      const int igt = grp_ids[iat] * ngroups;
      for (int j = 0; j < iat; ++j)
      {
        const BsplineFunctor<RT>& f = F[igt + grp_ids[j]];
        const RT r                  = row[j];
        RT uu(0), dd(0), d2(0);
        if (f.present && r < f.cutoff_radius)
        {
          // CPU functor path BsplineFunctor.h:772-838 evaluateVGL: grad = DeltaRInv * rinv * (...)
          uu = f.evaluate_impl(r, dd, d2);
          dd *= RT(1) / r;
        }
        u[j]   = uu;
        du[j]  = dd;
        d2u[j] = d2;
      }
      RT usum(0), lap(0), g[3] = {0, 0, 0};
      for (int j = 0; j < iat; ++j)
        usum += u[j];
      for (int j = 0; j < iat; ++j)
        lap += d2u[j] + RT(2) * du[j];
      for (int d = 0; d < 3; ++d)
      {
        RT sacc(0);
        for (int j = 0; j < iat; ++j)
          sacc += du[j] * row[(d + 1) * npad + j];
        g[d] = sacc;
      }
      s.Uat[iat]   = usum;
      dUx[iat]     = g[0];
      dUy[iat]     = g[1];
      dUz[iat]     = g[2];
      s.d2Uat[iat] = -lap;
      for (int j = 0; j < iat; ++j)
      {
        s.Uat[j] += u[j];
        s.d2Uat[j] -= d2u[j] + RT(2) * du[j];
        dUx[j] -= du[j] * row[npad + j];
        dUy[j] -= du[j] * row[2 * npad + j];
        dUz[j] -= du[j] * row[3 * npad + j];
      }
    }
    double lv = 0;
    for (int i = 0; i < N; ++i)
      lv -= 0.5 * (double)s.Uat[i];
    s.log_value = lv;
  }
};

// =====================================================================================
// One-body Jastrow (electron-ion).  ref: Jastrow/J1OrbitalSoA.h:136-185 (accumulateGL, computeU3),
// :441-469 (ratioGrad, acceptMove), :230-250 recompute; e-ion rows from SoaDistanceTableAB.h:59-70
// (d = R_ion - r_elec, flip_ind = 0).
template<typename RT>
struct OneBodyJastrow
{
  int Nions = 0, Nelec = 0, ngroups = 0;
  size_t npad_ion = 0;
  std::vector<RT> ion_rsoa; // [3][npad_ion]
  std::vector<int> ion_grp;
  std::vector<BsplineFunctor<RT>> F; // per ion group
  struct State
  {
    std::vector<RT> Vat, Grad /*[Nelec][3]*/, Lap;
    RT curAt, curLap, curGrad[3];
    double log_value = 0;
  };
  void init(int nions, const double* ion_pos, const int* igrp, int ng, int nelec)
  {
    Nions    = nions;
    Nelec    = nelec;
    ngroups  = ng;
    npad_ion = aligned_size<RT>(nions);
    ion_rsoa.assign(3 * npad_ion, RT(0));
    for (int i = 0; i < nions; ++i)
      for (int d = 0; d < 3; ++d)
        ion_rsoa[d * npad_ion + i] = (RT)ion_pos[3 * i + d];
    ion_grp.assign(igrp, igrp + nions);
    F.assign(ng, BsplineFunctor<RT>());
  }
  void initState(State& s) const
  {
    s.Vat.assign(Nelec, RT(0));
    s.Grad.assign(3 * (size_t)Nelec, RT(0));
    s.Lap.assign(Nelec, RT(0));
    s.log_value = 0;
  }
  void compute(const MinImage<RT>& mi, const RT pos[3], RT& at, RT& lap, RT grad[3]) const
  {
    RT usum(0), l(0), g0(0), g1(0), g2(0);
    std::vector<RT> u(Nions), du(Nions), d2u(Nions), dx(Nions), dy(Nions), dz(Nions);
    for (int j = 0; j < Nions; ++j)
    {
      RT r;
      mi.pair(pos, ion_rsoa[j], ion_rsoa[npad_ion + j], ion_rsoa[2 * npad_ion + j], j, 0, r, dx[j], dy[j], dz[j]);
      const BsplineFunctor<RT>& f = F[ion_grp[j]];
      RT uu(0), dd(0), d2(0);
      if (f.present && r < f.cutoff_radius)
      {
        uu = f.evaluate_impl(r, dd, d2);
        dd *= RT(1) / r;
      }
      u[j]   = uu;
      du[j]  = dd;
      d2u[j] = d2;
    }
    for (int j = 0; j < Nions; ++j)
      l += d2u[j] + RT(2) * du[j];
    for (int j = 0; j < Nions; ++j)
      g0 += du[j] * dx[j];
    for (int j = 0; j < Nions; ++j)
      g1 += du[j] * dy[j];
    for (int j = 0; j < Nions; ++j)
      g2 += du[j] * dz[j];
    for (int j = 0; j < Nions; ++j)
      usum += u[j];
    at      = usum;
    lap     = l;
    grad[0] = g0;
    grad[1] = g1;
    grad[2] = g2;
  }
  double ratioGrad(State& s, const MinImage<RT>& mi, int iat, const RT pos[3], RT grad_new[3]) const
  {
    compute(mi, pos, s.curAt, s.curLap, s.curGrad);
    for (int d = 0; d < 3; ++d)
      grad_new[d] += s.curGrad[d];
    return std::exp(static_cast<double>(s.Vat[iat] - s.curAt));
  }
  void acceptMove(State& s, int iat) const
  {
    s.log_value += s.Vat[iat] - s.curAt;
    s.Vat[iat] = s.curAt;
    for (int d = 0; d < 3; ++d)
      s.Grad[3 * iat + d] = s.curGrad[d];
    s.Lap[iat] = s.curLap;
  }
  // recompute: J1OrbitalSoA.h:237-250; log_value = -sum Vat (:122)
  void recompute(State& s, const MinImage<RT>& mi, const RT* rsoa, size_t npad_pos) const
  {
    double lv = 0;
    for (int iat = 0; iat < Nelec; ++iat)
    {
      RT pos[3] = {rsoa[iat], rsoa[npad_pos + iat], rsoa[2 * npad_pos + iat]};
      RT at, lap, g[3];
      compute(mi, pos, at, lap, g);
      s.Vat[iat] = at;
      s.Lap[iat] = lap;
      for (int d = 0; d < 3; ++d)
        s.Grad[3 * iat + d] = g[d];
      lv -= (double)at;
    }
    s.log_value = lv;
  }
};

// =====================================================================================
// RNG.  ref: Utilities/StdRandom.h:34-48 (uniform_real_distribution_as_boost on std::mt19937:
// (eng()-min)/(max-min+1) in double), Particle/ParticleBase/RandomSeqGenerator.h:33-52 (Box-Muller).
struct StdRandom
{
  std::mt19937 engine;
  explicit StdRandom(uint32_t seed = 911) : engine(seed) {}
  inline double operator()()
  {
    return static_cast<double>(engine() - engine.min()) / (static_cast<double>(engine.max() - engine.min()) + 1);
  }
};

template<typename T, typename RG>
inline void assignGaussRand(T* a, unsigned n, RG& rng)
{
  const double slightly_less_than_one = 1.0 - std::numeric_limits<double>::epsilon();
  const int nm1                       = n - 1;
  double temp1, temp2;
  for (int i = 0; i < nm1; i += 2)
  {
    temp1    = std::sqrt(-2.0 * std::log(1.0 - slightly_less_than_one * rng()));
    temp2    = 2.0 * M_PI * rng();
    a[i]     = temp1 * std::cos(temp2);
    a[i + 1] = temp1 * std::sin(temp2);
  }
  if (n % 2 == 1)
  {
    temp1  = std::sqrt(-2.0 * std::log(1.0 - slightly_less_than_one * rng()));
    temp2  = 2.0 * M_PI * rng();
    a[nm1] = temp1 * std::cos(temp2);
  }
}

// ref: QMCDrivers/GreenFunctionModifiers/DriftModifierUNR.cpp:20-31 (a_ = 1 default)
template<typename RT>
inline void getDrift(RT tau, const RT qf[3], RT drift[3], RT a = RT(1))
{
  drift[0]     = qf[0];
  drift[1]     = qf[1];
  drift[2]     = qf[2];
  const RT vsq = drift[0] * drift[0] + drift[1] * drift[1] + drift[2] * drift[2];
  const RT sc =
      vsq < std::numeric_limits<RT>::epsilon() ? tau : ((-1.0 + std::sqrt(1.0 + 2.0 * a * tau * vsq)) / (a * vsq));
  drift[0] *= sc;
  drift[1] *= sc;
  drift[2] *= sc;
}

} // namespace orc
