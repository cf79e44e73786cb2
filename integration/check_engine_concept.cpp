// integration/check_engine_concept.cpp -- compile-time proof that integration/DelayedUpdateB200.h is a drop-in for the
// update-engine concept of DiracDeterminantBatched (QMCWaveFunctions/Fermion/DiracDeterminantBatched.h:57-61).
//
// Never linked or run: `g++ -fsyntax-only` (integration/check.sh).  The function template below issues the engine calls
// exactly as DiracDeterminantBatched.cpp does (mw_evalGrad :209, mw_getInvRow :335, mw_accept_rejectRow :515,
// mw_updateInvMat :520,:562, mw_transferAinv_D2H :569, single-walker updateRow :453) with the argument types that file
// builds, and is instantiated for the reference's own DelayedUpdateBatched<OMPTARGET, VT> AND for DelayedUpdateB200<VT>:
// if either side changes a signature, a container alias or the resource type, this file stops compiling.
#include <complex>
#include <cstddef>
#include <memory>
#include <vector>
#include "config.h"
#include "OhmmsPETE/TinyVector.h"
#include "OhmmsPETE/OhmmsArray.h"
#include "OhmmsSoA/VectorSoaContainer.h"
#include "type_traits/template_types.hpp"
#include "OMPTarget/OffloadAlignedAllocators.hpp"
namespace qmcplusplus
{
struct QMCTraits // the one-line shim of SURVEY.md App. B (Configuration.h pulls libxml2, absent in this image)
{
  enum
  {
    DIM = 3
  };
};
} // namespace qmcplusplus
#include "QMCWaveFunctions/Fermion/DelayedUpdateBatched.h"
#include "DelayedUpdateB200.h"

namespace qmcplusplus
{
template<class UpdateEngine>
void drive_engine_like_DiracDeterminantBatched(int nw, int norb, int ndelay)
{
  using Value = typename UpdateEngine::Value;
  using Grad  = TinyVector<Value, 3>;
  using DualMatrix        = typename UpdateEngine::template DualMatrix<Value>;
  using OffloadMWVGLArray = typename UpdateEngine::template OffloadMWVGLArray<Value>;

  std::vector<std::unique_ptr<UpdateEngine>> engines_owned;
  std::vector<DualMatrix> psiMinv(nw);
  for (int iw = 0; iw < nw; ++iw)
  {
    engines_owned.push_back(std::make_unique<UpdateEngine>(norb, ndelay)); // det_engine_(NumOrbitals, ndelay) :87
    psiMinv[iw].resize(norb, norb);
  }
  RefVectorWithLeader<UpdateEngine> engine_list(*engines_owned[0]);
  RefVector<DualMatrix> psiMinv_refs;
  for (int iw = 0; iw < nw; ++iw)
  {
    engine_list.push_back(*engines_owned[iw]);
    psiMinv_refs.push_back(psiMinv[iw]);
  }
  typename UpdateEngine::MultiWalkerResource engine_rsc; // DiracDeterminantBatchedMultiWalkerResource::engine_rsc :73
  static_assert(std::is_same<decltype(engine_rsc.queue), compute::Queue<PlatformKind::OMPTARGET>>::value,
                "DiracDeterminantBatched enqueues its own transfers on engine_rsc.queue (:129, :581)");

  const int WorkingIndex = 0;
  std::vector<const Value*> dpsiM_row_list(nw, nullptr);
  std::vector<Grad> grad_now(nw);
  UpdateEngine::mw_evalGrad(engine_list, engine_rsc, psiMinv_refs, dpsiM_row_list, WorkingIndex, grad_now);

  std::vector<const Value*> rows = UpdateEngine::mw_getInvRow(engine_list, engine_rsc, psiMinv_refs, WorkingIndex, false);
  (void)rows;

  OffloadMWVGLArray phi_vgl_v;
  phi_vgl_v.resize(5, nw, norb);
  std::vector<Value*> psiM_g_dev_ptr_list(nw, nullptr), psiM_l_dev_ptr_list(nw, nullptr);
  std::vector<bool> isAccepted(nw, true);
  std::vector<Value> ratios_local(nw, Value(1));
  UpdateEngine::mw_accept_rejectRow(engine_list, engine_rsc, psiMinv_refs, WorkingIndex, psiM_g_dev_ptr_list,
                                    psiM_l_dev_ptr_list, isAccepted, phi_vgl_v, ratios_local);
  UpdateEngine::mw_updateInvMat(engine_list, engine_rsc, psiMinv_refs);
  UpdateEngine::mw_transferAinv_D2H(engine_list, engine_rsc, psiMinv_refs);

  Vector<Value> psiV(norb);
  engines_owned[0]->updateRow(psiMinv[0], WorkingIndex, psiV, Value(1));
}

// the reference's engine and the adapter, real and complex value types, full and mixed precision
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateBatched<PlatformKind::OMPTARGET, double>>(int, int, int);
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateB200<double>>(int, int, int);
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateBatched<PlatformKind::OMPTARGET, float>>(int, int, int);
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateB200<float>>(int, int, int);
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateBatched<PlatformKind::OMPTARGET, std::complex<double>>>(int, int, int);
template void drive_engine_like_DiracDeterminantBatched<DelayedUpdateB200<std::complex<double>>>(int, int, int);
} // namespace qmcplusplus
