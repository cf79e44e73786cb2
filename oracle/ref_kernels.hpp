// =====================================================================================
// oracle/ref_kernels.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Adapters that let the oracle driver run on the REFERENCE's own compiled kernels.  The reference
// headers are included from where they lie under /root/reference/src (see oracle/Makefile `ref`);
// no reference source is copied into this repository.
//   spline2::evaluate_v_impl / evaluate_vgl_impl / evaluate_vgh_impl   (spline2/MultiBsplineEval.hpp)
//   qmcplusplus::DelayedUpdate<T>                                     (QMCWaveFunctions/Fermion/DelayedUpdate.h)
//   qmcplusplus::DiracMatrix<T>                                       (QMCWaveFunctions/Fermion/DiracMatrix.h)
// =====================================================================================
#pragma once
#include "config.h"
#include <vector>
#include <cstddef>
#include <complex>
#include "CPU/SIMD/aligned_allocator.hpp"
#include "spline2/MultiBsplineEval.hpp"
#include "QMCWaveFunctions/Fermion/DelayedUpdate.h"
#include "QMCWaveFunctions/Fermion/DiracMatrix.h"
#include "qmc_oracle.hpp"

namespace orc
{
template<typename ST>
struct RefSplineType;
template<>
struct RefSplineType<float>
{
  using type = multi_UBspline_3d_s;
};
template<>
struct RefSplineType<double>
{
  using type = multi_UBspline_3d_d;
};

template<typename ST>
inline typename RefSplineType<ST>::type make_ref_spline(const SplineTable<ST>& s)
{
  typename RefSplineType<ST>::type m;
  std::memset(&m, 0, sizeof(m));
  m.coefs    = const_cast<ST*>(s.coefs);
  m.x_stride = s.xs;
  m.y_stride = s.ys;
  m.z_stride = s.zs;
  Ugrid* grids[3] = {&m.x_grid, &m.y_grid, &m.z_grid};
  for (int d = 0; d < 3; ++d)
  {
    grids[d]->start     = 0.0;
    grids[d]->end       = 1.0;
    grids[d]->num       = s.M[d];
    grids[d]->delta     = 1.0 / s.M[d];
    grids[d]->delta_inv = s.delta_inv[d];
  }
  m.num_splines = s.ns;
  m.coefs_size  = s.coefs_size();
  return m;
}

template<typename ST>
inline void kernel_vgh(const SplineTable<ST>& s, ST x, ST y, ST z, ST* v, ST* g, ST* h, size_t stride)
{
  auto m = make_ref_spline(s);
  spline2::evaluate_vgh_impl(&m, x, y, z, v, g, h, stride, 0, s.ns);
}

// the reference's CPU engine assumes lda == norb (DelayedUpdate.h:93,120: Ainv.data(), norb)
template<typename T>
struct Engine
{
  qmcplusplus::DelayedUpdate<T> eng;
  int n = 0, k = 0;
  int delay_count = 0;
  void resize(int norb, int delay)
  {
    n = norb;
    k = delay;
    eng.resize(norb, delay);
    delay_count = 0;
  }
  void getInvRow(const T* Ainv, int lda, int row, T* invRow)
  {
    if (lda != n)
      throw std::runtime_error("reference DelayedUpdate needs lda == norb");
    qmcplusplus::Matrix<T> A(const_cast<T*>(Ainv), n, n);
    qmcplusplus::Vector<T> r(invRow, n);
    eng.getInvRow(A, row, r);
  }
  template<typename RATIOT>
  void acceptRow(T* Ainv, int lda, int row, const T* psiV, const RATIOT ratio)
  {
    qmcplusplus::Matrix<T> A(Ainv, n, n);
    qmcplusplus::Vector<T> v(const_cast<T*>(psiV), n);
    eng.acceptRow(A, row, v, ratio);
    delay_count = eng.getDelayCount();
  }
  void pseudoAcceptRow(T*, int, int) { throw std::runtime_error("reference CPU DelayedUpdate has no pseudo-accept"); }
  void updateInvMat(T* Ainv, int lda)
  {
    qmcplusplus::Matrix<T> A(Ainv, n, n);
    eng.updateInvMat(A);
    delay_count = 0;
  }
};

template<typename VT>
inline void kernel_invert_transpose(const VT* amat, int n, int a_cols, VT* inv, int lda, std::complex<double>& logdet)
{
  if (a_cols != n || lda != n)
    throw std::runtime_error("reference invert_transpose adapter needs unpadded matrices");
  qmcplusplus::DiracMatrix<typename FullPrec<VT>::type> dm;
  qmcplusplus::Matrix<VT> a(const_cast<VT*>(amat), n, n);
  qmcplusplus::Matrix<VT> b(inv, n, n);
  dm.invert_transpose(a, b, logdet);
}
} // namespace orc
