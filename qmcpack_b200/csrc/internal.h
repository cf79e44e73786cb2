// qmcpack_b200/csrc/internal.h -- host-side classes behind the C ABI (include/qmcb.h).
// Names mirror the reference classes they stand in for (SplineR2R/SplineC2C, DiracDeterminantBatched +
// DelayedUpdateBatched, TwoBodyJastrow, SoaDistanceTableAA, J1OrbitalSoA, VMCBatched) so that the in-tree adapter of
// INTEGRATION.md is a thin forwarder.
#pragma once
#include "common.cuh"
#include "../../include/qmcb.h"
#include <vector>
#include <memory>

namespace qmcb
{
// ------------------------------------------------------------------------------------------------
// SPOSet (read-only table in HBM; evaluation scratch owned per call site)
struct SplineSPOBase
{
  int device = 0; // CUDA device the table lives on (the current device at creation)
  int precision = 0, kind = 0;
  int grid[3]   = {0, 0, 0};
  int n_orb = 0, n_spl = 0;
  size_t npad = 0;
  double G[9];
  int halfG[3] = {0, 0, 0};
  virtual ~SplineSPOBase() {}
  virtual size_t table_bytes() const = 0;
  int vt_per_orb() const { return kind == QMCB_C2C ? 2 : 1; } // scalars per orbital value
  size_t elem_size() const { return precision == QMCB_MIXED ? 4 : 8; }
  // number of partial ratio/gradient slots per walker written by one evaluation (tiles x consumer warps)
  virtual int rg_parts() const = 0;
  // device-resident evaluation.  r_dev [nw][3] RT; invrow_dev [*][ld_inv] (may be null); ref_dev optional row map;
  // phi_dev: MODE_VGL [5][nw][n_orb], MODE_V [nw][n_orb] (may be null); rg_dev [nw][rg_parts()][4*vt] partial
  // dots to be added in index order (may be null)
  virtual void evaluate_dev(int mode, int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv,
                            const int* ref_dev, void* phi_dev, void* rg_dev, cudaStream_t st) = 0;
  // for the fused walker-segment kernel (segment.cuh): the device-side table descriptor (SplineDev<ST>) and the tensor
  // map whose box is one 8-row slab of the stencil (192 components x 4 z x 2 y x 1 x); CUtensorMap*, R2R tables only
  virtual const void* dev_desc() const       = 0;
  virtual const void* seg_tensor_map() const = 0;
};

SplineSPOBase* make_spline(int precision, int kind, const int grid[3], int n_orb, int n_spl, size_t npad,
                           const void* coefs_host, const double G[9], const int halfG[3], const double* kcart);

// ------------------------------------------------------------------------------------------------
// crowd = nw walkers with all wavefunction state; typed implementation behind a virtual interface
struct CrowdBase
{
  int device = 0; // CUDA device of this crowd; every ABI entry binds the calling host thread to it (one crowd per
                  // host thread, and new threads start on device 0)
  virtual ~CrowdBase() {}
  virtual void sync()                                                                                       = 0;
  virtual size_t device_bytes() const                                                                       = 0;
  virtual bool is_complex() const                                                                           = 0;
  virtual void set_positions(const double* R)                                                               = 0;
  virtual void get_positions(double* R)                                                                     = 0;
  virtual void twf_recompute()                                                                              = 0;
  virtual void twf_eval_grad(int iat, double* grads)                                                        = 0;
  virtual void ps_make_move(int iat, const double* displ)                                                   = 0;
  virtual void twf_calc_ratio_grad(int iat, double* ratios, double* grads)                                  = 0;
  virtual void twf_accept_reject(int iat, const uint8_t* acc, int safe_to_delay)                            = 0;
  virtual void twf_complete_updates()                                                                       = 0;
  virtual void twf_evaluate_ratios(int nvp, const int* walker, const int* ref, const double* r_vp, int ct, double* ratios) = 0;
  virtual void twf_calc_ratio(int iat, double* ratios)                                                      = 0;
  virtual void twf_evaluate_gl(double* G, double* L, double* logpsi, double* ke)                            = 0;
  virtual void det_eval_grad(int spin, int row, void* grads)                                                = 0;
  virtual void det_get_inv_row(int spin, int row, const void** dev, size_t* ld, void* host)                 = 0;
  virtual void det_ratio_grad(int spin, int row, void* ratios, void* grads, bool from_phi)                  = 0;
  virtual void det_accept_reject(int spin, int row, const uint8_t* acc)                                     = 0;
  virtual void det_complete_updates(int spin, void* psiMinv, double* logdet)                                = 0;
  virtual void det_recompute_from_matrices(int spin, const void* psiM, const void* dpsiM, const void* d2psiM) = 0;
  virtual void det_set_phi_vgl(int spin, const void* phi)                                                   = 0;
  virtual int det_delay_count(int spin)                                                                     = 0;
  virtual void det_time_update_inv_mat(int spin, int c, int reps, double* us_per_call)                      = 0;
  virtual void det_time_inverse(int spin, int method, int reps, double* us_per_call)                       = 0;
  virtual void dtaa_get_temp_rows(void* rows)                                                               = 0;
  virtual void j2_ratio_grad(int iat, double* ratios, void* grads)                                          = 0;
  virtual void j2_accept_reject(int iat, const uint8_t* acc)                                                = 0;
  virtual void j2_get_state(int iw, double* Uat, double* dUat, double* d2Uat)                               = 0;
  virtual void vmc_init(const qmcb_vmc_params* p)                                                           = 0;
  virtual void vmc_sweep(int nsteps, uint8_t* accept_log)                                                   = 0;
  virtual void vmc_sweep_async()                                                                            = 0;
  virtual void vmc_counts(long long* na, long long* nr)                                                     = 0;
  virtual int vmc_sweep_kernel() const                                                                      = 0;
  virtual int host_kernel() const                                                                           = 0;
  virtual void vmc_profile_sweep(double* out9)                                                              = 0;
  virtual void dmc_get_rr(double* rr_acc, double* rr_prop)                                                  = 0;
  virtual size_t walker_bytes() const                                                                       = 0;
  virtual void pack_walker(int iw, void* dev_buf)                                                           = 0;
  virtual void unpack_walker(int iw, const void* dev_buf)                                                   = 0;
  virtual void copy_walker(int src, int dst)                                                                = 0;
  virtual void set_num_walkers(int n_active)                                                                = 0;
  virtual int num_walkers() const                                                                           = 0;
  virtual int capacity() const                                                                              = 0;
  virtual cudaStream_t stream()                                                                             = 0;
};

CrowdBase* make_crowd(const qmcb_system* sys, int nw);

} // namespace qmcb

struct qmcb_spline
{
  std::unique_ptr<qmcb::SplineSPOBase> impl;
};
struct qmcb_crowd
{
  std::unique_ptr<qmcb::CrowdBase> impl;
};
