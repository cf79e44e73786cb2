// integration/SplineB200.h -- the SPOSet side of the binding: what `class SplineB200 : public BsplineSet` forwards to.
//
// SPOSet.h / BsplineSet.h cannot be compiled in this image (Configuration.h pulls libxml2), so the
// adapter is split in two: `SplineB200Core<ST, VT>` below holds every line that touches the C ABI and is written against
// the reference's own containers and einspline structs (compile-checked by integration/check_spline_adapter.cpp); the
// derived class a maintainer adds in QMCWaveFunctions/BsplineFactory/ is the ten forwarding lines at the end of this file
// (selection point: SplineSetReader.cpp:98-119, next to SplineR2R / SplineC2COMPTarget).
//
// Contract kept from the reference (SPOSet.h:346-352, DiracDeterminantBatched.cpp:334-346): isOMPoffload() is true, so the
// determinant passes DEVICE inverse rows and expects the device copy of phi_vgl_v to be current on return, ratios and
// gradients on the host.  With integration/DelayedUpdateB200.h as the update engine the rows are one [nw][ld] block
// (row iw = base + iw * ld); the adapter checks that instead of assuming it.
#ifndef QMCPLUSPLUS_SPLINE_B200_H
#define QMCPLUSPLUS_SPLINE_B200_H

#include <complex>
#include <memory>
#include <stdexcept>
#include <vector>
#include "OhmmsPETE/TinyVector.h"
#include "OhmmsPETE/Tensor.h"
#include "OhmmsPETE/OhmmsVector.h"
#include "OhmmsPETE/OhmmsArray.h"
#include "OMPTarget/OffloadAlignedAllocators.hpp"
#include "spline2/bspline_traits.hpp"
#include "type_traits/complex_help.hpp"
#include "qmcb.h"

namespace qmcplusplus
{
template<typename ST, typename VT>
class SplineB200Core
{
public:
  using SplineType        = typename bspline_traits<ST, 3>::SplineType; // multi_UBspline_3d_{s,d}
  using ValueType         = VT;
  using GradType          = TinyVector<VT, 3>;
  using PosType           = TinyVector<double, 3>;
  using ValueVector       = Vector<VT>;
  using GradVector        = Vector<GradType>;
  using OffloadMWVGLArray = Array<VT, 3, OffloadPinnedAllocator<VT>>; // [VGL, walker, Orbs]
  static constexpr bool is_complex = IsComplex_t<VT>::value;
  static_assert(std::is_same<RealAlias<VT>, ST>::value, "orbital value type and spline storage share one precision");

  /// SPOSet::finalizeConstruction (SPOSet.h:577): the host table as SplineSetReader leaves it goes to HBM once.
  /// PrimLattice.G / HalfG / kPoints are the members of BsplineSet (BsplineSet.h:38-252); the SIMULATION cell may be a
  /// tiling, the spline lives on the primitive cell.
  void finalizeConstruction(const SplineType& spline,
                            int orbital_set_size,
                            const Tensor<double, 3>& prim_G,
                            const TinyVector<int, 3>& half_g,
                            const std::vector<PosType>& k_points)
  {
    const int grid[3] = {spline.x_grid.num, spline.y_grid.num, spline.z_grid.num};
    const int hg[3]   = {half_g[0], half_g[1], half_g[2]};
    double G[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        G[3 * i + j] = prim_G(i, j);
    std::vector<double> kc;
    if (is_complex)
      for (int j = 0; j < orbital_set_size; ++j)
        for (int d = 0; d < 3; ++d)
          kc.push_back(k_points[j][d]);
    qmcb_spline* h = nullptr;
    check(qmcb_spline_create(&h, sizeof(ST) == 4 ? QMCB_MIXED : QMCB_FULL, is_complex ? QMCB_C2C : QMCB_R2R, grid,
                             orbital_set_size, (is_complex ? 2 : 1) * orbital_set_size, spline.z_stride, spline.coefs, G, hg,
                             is_complex ? kc.data() : nullptr));
    table_.reset(h, qmcb_spline_destroy); // clones share the table like SplineR2R::SplineInst (SplineR2R.h:82)
    norb_ = orbital_set_size;
  }

  bool isOMPoffload() const { return true; }
  qmcb_spline* handle() const { return table_.get(); }

  /// SPOSet::mw_evaluateValue (SPOSet.h:300) with the positions already gathered from P_list[iw].activeR(iat)
  void mw_evaluateValue(const std::vector<PosType>& pos, const RefVector<ValueVector>& psi_v_list) const
  {
    const int nw = pos.size();
    stage_.resize((size_t)nw * norb_);
    check(qmcb_spline_mw_evaluate_value(table_.get(), nw, &pos[0][0], stage_.data()));
    for (int iw = 0; iw < nw; ++iw)
      std::copy_n(stage_.data() + (size_t)iw * norb_, norb_, psi_v_list[iw].get().data());
  }

  /// SPOSet::mw_evaluateVGL (SPOSet.h:313)
  void mw_evaluateVGL(const std::vector<PosType>& pos,
                      const RefVector<ValueVector>& psi_v_list,
                      const RefVector<GradVector>& dpsi_v_list,
                      const RefVector<ValueVector>& d2psi_v_list) const
  {
    const int nw = pos.size();
    stage_.resize((size_t)nw * norb_ * 5);
    VT* psi  = stage_.data();
    VT* dpsi = psi + (size_t)nw * norb_;
    VT* d2   = dpsi + (size_t)nw * norb_ * 3;
    check(qmcb_spline_mw_evaluate_vgl(table_.get(), nw, &pos[0][0], psi, dpsi, d2));
    for (int iw = 0; iw < nw; ++iw)
    {
      std::copy_n(psi + (size_t)iw * norb_, norb_, psi_v_list[iw].get().data());
      std::copy_n(d2 + (size_t)iw * norb_, norb_, d2psi_v_list[iw].get().data());
      for (int j = 0; j < norb_; ++j)
        for (int d = 0; d < 3; ++d)
          dpsi_v_list[iw].get()[j][d] = dpsi[((size_t)iw * norb_ + j) * 3 + d];
    }
  }

  /// SPOSet::mw_evaluateVGLandDetRatioGrads (SPOSet.h:346-352).  `queue_native` = the crowd's stream
  /// (compute::Queue<PL>::getNative(), Platforms/CUDA/QueueCUDA.hpp:27-64)
  void mw_evaluateVGLandDetRatioGrads(const std::vector<PosType>& pos,
                                      const std::vector<const VT*>& invRow_ptr_list,
                                      OffloadMWVGLArray& phi_vgl_v,
                                      std::vector<VT>& ratios,
                                      std::vector<GradType>& grads,
                                      void* queue_native) const
  {
    const size_t nw = pos.size();
    if (nw == 0)
      return;
    // rows handed out by DelayedUpdateB200::mw_getInvRow are equidistant; anything else is a caller this adapter
    // does not serve
    const std::ptrdiff_t ld = nw > 1 ? invRow_ptr_list[1] - invRow_ptr_list[0] : norb_;
    for (size_t iw = 1; iw < nw; ++iw)
      if (invRow_ptr_list[iw] - invRow_ptr_list[0] != (std::ptrdiff_t)iw * ld)
        throw std::runtime_error("SplineB200: inverse rows are not one equidistant device block");
    if (ld < norb_)
      throw std::runtime_error("SplineB200: inverse-row stride smaller than the orbital count");
    check(qmcb_spline_mw_evaluate_vgl_ratio_grads_offload(table_.get(), (int)nw, &pos[0][0], invRow_ptr_list[0], (size_t)ld,
                                                          phi_vgl_v.device_data(), ratios.data(), &grads[0][0],
                                                          queue_native));
  }

  /// SPOSet::mw_evaluateDetRatios (SPOSet.h:257) for the virtual particles of the non-local pseudopotential: positions
  /// and reference-walker map gathered from the VirtualParticleSets, host inverse rows [n_ref][ld]
  void mw_evaluateDetRatios(const std::vector<PosType>& vp_pos,
                            const std::vector<int>& ref_walker,
                            int n_ref,
                            const VT* inv_rows_host,
                            size_t ld,
                            std::vector<VT>& ratios) const
  {
    ratios.resize(vp_pos.size());
    if (!vp_pos.empty())
      check(qmcb_spline_mw_evaluate_det_ratios(table_.get(), (int)vp_pos.size(), &vp_pos[0][0], ref_walker.data(), n_ref,
                                               inv_rows_host, ld, ratios.data()));
  }

private:
  static void check(int rc)
  {
    if (rc != 0)
      throw std::runtime_error(qmcb_last_error());
  }
  std::shared_ptr<qmcb_spline> table_;
  int norb_ = 0;
  mutable std::vector<VT> stage_;
};

/* The derived class (not compilable here: BsplineSet.h -> SPOSet.h -> Configuration.h -> libxml2):

template<typename ST>
class SplineB200 : public BsplineSet
{
  SplineB200Core<ST, ValueType> core_;
  std::shared_ptr<MultiBspline<ST>> SplineInst;   // filled by SplineSetReader exactly as for SplineR2R
public:
  std::string getClassName() const override { return "SplineB200"; }
  bool isOMPoffload() const override { return true; }
  void finalizeConstruction() override
  { core_.finalizeConstruction(*SplineInst->getSplinePtr(), OrbitalSetSize, PrimLattice.G, HalfG, kPoints); }
  void mw_evaluateVGLandDetRatioGrads(const RefVectorWithLeader<SPOSet>& spo_list, const RefVectorWithLeader<ParticleSet>& P_list,
                                      int iat, const std::vector<const ValueType*>& invRow_ptr_list, OffloadMWVGLArray& phi_vgl_v,
                                      std::vector<ValueType>& ratios, std::vector<GradType>& grads) const override
  {
    std::vector<PosType> pos; for (const ParticleSet& P : P_list) pos.push_back(P.activeR(iat));
    core_.mw_evaluateVGLandDetRatioGrads(pos, invRow_ptr_list, phi_vgl_v, ratios, grads, queue.getNative());
  }
  // mw_evaluateValue / mw_evaluateVGL / mw_evaluateDetRatios gather positions the same way and forward.
};
*/
} // namespace qmcplusplus
#endif
