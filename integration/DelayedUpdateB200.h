// integration/DelayedUpdateB200.h -- the determinant update engine a QMCPACK maintainer adds to bind libqmcb.so.
//
// `DiracDeterminantBatched<PL, VT, FPVT>` takes its update engine from `UpdateEngineSelector<PL, VT>::Engine`
// (QMCWaveFunctions/Fermion/DiracDeterminantBatched.h:32-61) and only ever talks to it through the static mw_* calls and
// the nested MultiWalkerResource of `DelayedUpdateBatched<PL, VALUE>` (Fermion/DelayedUpdateBatched.h:35-830).  This class
// has the same member names, argument lists and container types, and forwards each call to the C ABI of include/qmcb.h.
// It is compiled (syntax-only, against the reference's own headers) together with the reference's engine by
// integration/check_engine_concept.cpp: one function template drives BOTH engines through the call sequence of
// DiracDeterminantBatched.cpp, so a signature drift on either side breaks the check (integration/check.sh,
// tests/test_integration_cpu.py).
//
// Ownership.  The reference keeps psiMinv in each walker's DiracDeterminantBatched (a DualMatrix, host + device copy) and
// the delay buffers U, V, Binv in each walker's engine object.  libqmcb keeps all of it in HBM inside one `qmcb_crowd`
// per crowd (walker-major blocks, include/qmcb.h), so the per-walker engine objects of this adapter are empty shells and
// the MultiWalkerResource carries the crowd handle (created in `createResource`, lent with the rest of the crowd's
// resources, Utilities/ResourceCollection.h; VMCBatched.cpp:79-80) and the index of the spin determinant it serves.
// psiMinv_refs is honoured where the reference hands data to the caller: mw_transferAinv_D2H fills the host copies.
//
// Error convention: every qmcb_* call returns non-zero on failure; the adapter throws std::runtime_error with
// qmcb_last_error(), the reference's convention (DelayedUpdateBatched.h:164-168, DiracDeterminantBatched.cpp:494-500).
#ifndef QMCPLUSPLUS_DELAYED_UPDATE_B200_H
#define QMCPLUSPLUS_DELAYED_UPDATE_B200_H

#include <complex>
#include <cstdint>
#include <stdexcept>
#include <vector>
#include "OhmmsPETE/OhmmsVector.h"
#include "OhmmsPETE/OhmmsMatrix.h"
#include "OhmmsPETE/OhmmsArray.h"
#include "OMPTarget/OffloadAlignedAllocators.hpp"
#include "QueueAliases.hpp"
#include "type_traits/complex_help.hpp"
#include "type_traits/template_types.hpp"
#include "qmcb.h"

namespace qmcplusplus
{
#define QMCB_CHECK(call)                               \
  do                                                   \
  {                                                    \
    if ((call) != 0)                                   \
      throw std::runtime_error(qmcb_last_error());     \
  } while (0)

template<typename VALUE>
class DelayedUpdateB200
{
public:
  using This_t  = DelayedUpdateB200<VALUE>;
  using Value   = VALUE;
  using Real    = RealAlias<Value>;
  using Complex = std::complex<Real>;

  // the container aliases DiracDeterminantBatched names through its engine (DelayedUpdateBatched.h:45-56)
  template<typename DT>
  using UnpinnedDualVector = Vector<DT, OffloadAllocator<DT>>;
  template<typename DT>
  using DualVector = Vector<DT, OffloadPinnedAllocator<DT>>;
  template<typename DT>
  using DualMatrix = Matrix<DT, OffloadPinnedAllocator<DT>>;
  template<typename DT>
  using OffloadMWVGLArray = Array<DT, 3, OffloadPinnedAllocator<DT>>; // [VGL, walker, Orbs]
  template<typename DT>
  using OffloadMatrix = Matrix<DT, OffloadPinnedAllocator<DT>>;

  /// per-crowd resource (DelayedUpdateBatched.h:58-109): the reference's stream + BLAS handle + pointer buffers become
  /// one crowd handle; `queue` stays because DiracDeterminantBatched enqueues its own transfers on it
  /// (DiracDeterminantBatched.cpp:129,581)
  struct MultiWalkerResource
  {
    compute::Queue<PlatformKind::OMPTARGET> queue;
    qmcb_crowd* crowd = nullptr; ///< borrowed; owns walker state, delay buffers and the CUDA stream of this crowd
    int spin          = 0;       ///< which determinant of the crowd this component is
    std::vector<uint8_t> flags;  ///< isAccepted as bytes
    std::vector<Value> grads;    ///< [nw][3] gradients in the value type
    std::vector<Value> ainv;     ///< [nw][n][n] staging block of mw_transferAinv_D2H
    void resize_fill_constant_arrays(size_t) {}
  };

  /// DelayedUpdateBatched(size_t norb, size_t max_delay): the delay rank is a property of the crowd
  /// (qmcb_system::delay_rank); kept for the checks of DiracDeterminantBatched
  DelayedUpdateB200(size_t norb, size_t max_delay) : norb_(norb), delay_(max_delay) {}
  DelayedUpdateB200(const DelayedUpdateB200&) = delete;

  /// DelayedUpdateBatched::mw_evalGrad (:354-400): prepares the inverse rows incl. pending delays and returns
  /// grad_now = invRow . dpsiM[row].  dpsiM_row_list is unused: the gradient rows live in the crowd (saved by the
  /// accepts, DelayedUpdateBatched.h:640-660)
  template<typename GT>
  static void mw_evalGrad(const RefVectorWithLeader<This_t>& engines,
                          MultiWalkerResource& mw_rsc,
                          const RefVector<DualMatrix<Value>>& psiMinv_refs,
                          const std::vector<const Value*>& dpsiM_row_list,
                          const int rowchanged,
                          std::vector<GT>& grad_now)
  {
    const size_t nw = engines.size();
    mw_rsc.grads.resize(3 * nw);
    QMCB_CHECK(qmcb_det_mw_eval_grad(mw_rsc.crowd, mw_rsc.spin, rowchanged, mw_rsc.grads.data()));
    for (size_t iw = 0; iw < nw; ++iw)
      grad_now[iw] = {mw_rsc.grads[3 * iw], mw_rsc.grads[3 * iw + 1], mw_rsc.grads[3 * iw + 2]};
  }

  /// spinor wavefunctions are outside the path this library serves
  template<typename GT>
  static void mw_evalGradWithSpin(const RefVectorWithLeader<This_t>&,
                                  MultiWalkerResource&,
                                  const RefVector<DualMatrix<Value>>&,
                                  const std::vector<const Value*>&,
                                  OffloadMatrix<Complex>&,
                                  const int,
                                  std::vector<GT>&,
                                  std::vector<Complex>&)
  {
    throw std::runtime_error("DelayedUpdateB200: spinor gradients are not served by libqmcb");
  }

  /// single-walker Sherman-Morrison update (DelayedUpdateBatched.h:492-536): the legacy per-walker API is not served
  template<typename VVT, typename FPVT>
  void updateRow(DualMatrix<Value>&, int, const VVT&, FPVT)
  {
    throw std::runtime_error("DelayedUpdateB200: single-walker updateRow is not served; use the mw_ API");
  }

  /// DelayedUpdateBatched::mw_accept_rejectRow (:542-670): accepted walkers append (U, V, bordered Binv update, G/L rows),
  /// rejected walkers pseudo-accept; the orbital rows of the proposed move (phi_vgl_v) are already in the crowd -- the
  /// gather that produced them wrote them there (SPOSet::mw_evaluateVGLandDetRatioGrads, SPOSet.h:346-352)
  static void mw_accept_rejectRow(const RefVectorWithLeader<This_t>& engines,
                                  MultiWalkerResource& mw_rsc,
                                  const RefVector<DualMatrix<Value>>& psiMinv_refs,
                                  const int rowchanged,
                                  const std::vector<Value*>& psiM_g_list,
                                  const std::vector<Value*>& psiM_l_list,
                                  const std::vector<bool>& isAccepted,
                                  const OffloadMWVGLArray<Value>& phi_vgl_v,
                                  const std::vector<Value>& ratios)
  {
    const size_t nw = engines.size();
    mw_rsc.flags.resize(nw);
    for (size_t iw = 0; iw < nw; ++iw)
      mw_rsc.flags[iw] = isAccepted[iw] ? 1 : 0;
    QMCB_CHECK(qmcb_det_mw_accept_reject(mw_rsc.crowd, mw_rsc.spin, rowchanged, mw_rsc.flags.data()));
  }

  /// DelayedUpdateBatched::mw_updateInvMat (:675-738): the rank-k Woodbury flush
  static void mw_updateInvMat(const RefVectorWithLeader<This_t>& engines,
                              MultiWalkerResource& mw_rsc,
                              const RefVector<DualMatrix<Value>>& psiMinv_refs)
  {
    QMCB_CHECK(qmcb_det_mw_complete_updates(mw_rsc.crowd, mw_rsc.spin, nullptr, nullptr));
  }

  /// DelayedUpdateBatched::mw_getInvRow (:763-810): rows of the current inverse incl. pending delays, one pointer per
  /// walker; device pointers into the crowd's contiguous [nw][ld] block, or host pointers into the engine's staging
  static std::vector<const Value*> mw_getInvRow(const RefVectorWithLeader<This_t>& engines,
                                                MultiWalkerResource& mw_rsc,
                                                const RefVector<DualMatrix<Value>>& psiMinv_refs,
                                                const int row_id,
                                                bool on_host)
  {
    const size_t nw   = engines.size();
    const size_t norb = engines.getLeader().norb_;
    const void* dev   = nullptr;
    size_t ld         = 0;
    std::vector<const Value*> row_ptr_list;
    row_ptr_list.reserve(nw);
    if (on_host)
    {
      mw_rsc.ainv.resize(nw * norb);
      QMCB_CHECK(qmcb_det_mw_get_inv_row(mw_rsc.crowd, mw_rsc.spin, row_id, &dev, &ld, mw_rsc.ainv.data()));
      for (size_t iw = 0; iw < nw; ++iw)
        row_ptr_list.push_back(mw_rsc.ainv.data() + iw * norb);
    }
    else
    {
      QMCB_CHECK(qmcb_det_mw_get_inv_row(mw_rsc.crowd, mw_rsc.spin, row_id, &dev, &ld, nullptr));
      for (size_t iw = 0; iw < nw; ++iw)
        row_ptr_list.push_back(static_cast<const Value*>(dev) + iw * ld);
    }
    return row_ptr_list;
  }

  /// DelayedUpdateBatched::mw_transferAinv_D2H (:813-826): the host copies of psiMinv become current (padding of the
  /// reference's rows, psiMinv.cols() >= norb, is left untouched)
  static void mw_transferAinv_D2H(const RefVectorWithLeader<This_t>& engines,
                                  MultiWalkerResource& mw_rsc,
                                  const RefVector<DualMatrix<Value>>& psiMinv_refs)
  {
    const size_t nw   = engines.size();
    const size_t norb = engines.getLeader().norb_;
    mw_rsc.ainv.resize(nw * norb * norb);
    QMCB_CHECK(qmcb_det_mw_complete_updates(mw_rsc.crowd, mw_rsc.spin, mw_rsc.ainv.data(), nullptr));
    for (size_t iw = 0; iw < nw; ++iw)
    {
      DualMatrix<Value>& psiMinv = psiMinv_refs[iw];
      for (size_t i = 0; i < norb; ++i)
        std::copy_n(mw_rsc.ainv.data() + (iw * norb + i) * norb, norb, psiMinv[i]);
    }
  }

  size_t norb() const { return norb_; }
  size_t delay() const { return delay_; }

private:
  size_t norb_, delay_;
};

#undef QMCB_CHECK
} // namespace qmcplusplus
#endif
