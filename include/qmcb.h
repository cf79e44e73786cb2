/* =====================================================================================================
 * include/qmcb.h -- C ABI of libqmcb.so, the B200 (sm_100a) batched per-electron-move engine.
 *
 * Drop-in boundary for QMCPACK's batched drivers.  QMCPACK has no C ABI / plugin loader for this path: the
 * path sits behind C++ virtual interfaces.  Each entry point below names the reference interface it replaces
 * (paths relative to /root/reference/src); INTEGRATION.md shows the thin in-tree adapter classes
 * (class SplineB200 : public BsplineSet, an UpdateEngine for DiracDeterminantBatched, a TwoBodyJastrow
 * forwarding class) a maintainer would add to bind them.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; qmcb_last_error() (thread-local) gives the message.
 *    The adapter turns non-zero into `throw std::runtime_error(qmcb_last_error())`, the reference's error
 *    convention (DiracDeterminantBatched.cpp:443-448, DelayedUpdateBatched.h:164-168).
 *  - plain pointers and sizes only.  `_host` arguments are host buffers (pinned or pageable), copied inside
 *    the call; `_dev` arguments are device pointers that stay resident (the reference passes device pointers
 *    for the inverse rows when SPOSet::isOMPoffload() is true, DiracDeterminantBatched.cpp:334-336).
 *  - precision: QMCB_FULL = RealType/ValueType/spline all double; QMCB_MIXED = all float with the matrix
 *    inversion and log-determinants in double (the reference's MIXED_PRECISION build).
 *  - handles are thread-compatible; one qmcb_crowd (walker batch + stream) per host thread, like one Crowd /
 *    MultiWalkerResource per OpenMP thread in the reference (VMCBatched.cpp:348,405; DelayedUpdateBatched.h:58-109).
 *  - there is NO CPU fallback: every call fails with a clear error when no CUDA device is present.
 * ===================================================================================================== */
#ifndef QMCB_H
#define QMCB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QMCB_FULL 0
#define QMCB_MIXED 1
#define QMCB_R2R 0 /* SplineR2R: real table -> real orbitals      (BsplineFactory/SplineR2R.h:36)  */
#define QMCB_C2C 1 /* SplineC2C: real pairs + twist -> complex    (BsplineFactory/SplineC2C.h:36)  */

typedef struct qmcb_spline qmcb_spline; /* one SPOSet (read-only table, shared by clones like SplineR2R.h:82) */
typedef struct qmcb_crowd qmcb_crowd;   /* nw walkers: positions, 2 determinants, J1/J2, scratch, one stream   */

/* ---- library ------------------------------------------------------------------------------------ */
int qmcb_init(int device);             /* cudaSetDevice + capability check (sm_100 required)            */
const char* qmcb_last_error(void);
int qmcb_device_count(void);           /* 0 when no usable CUDA device (never an error)                  */
size_t qmcb_aligned_size(int precision, size_t n); /* getAlignedSize<T> (Platforms/CPU/SIMD/aligned_allocator.hpp:41) */
/* number of kernel launches issued by this library since load (bench.py's gpu_launches counter) */
unsigned long long qmcb_kernel_launch_count(void);

/* ---- SPOSet: tricubic B-spline orbitals ----------------------------------------------------------- *
 * replaces BsplineSet / SplineR2R / SplineC2C(OMPTarget) evaluation entry points.                     */
/* coefs_host: [grid[0]+3][grid[1]+3][grid[2]+3][npad] in ST (multi_UBspline_3d_{s,d} layout,
 * spline2/MultiBsplineBase.hpp:75-128).  n_spl = real components (n_orb for R2R, 2*n_orb for C2C).
 * G: prim_lattice.G row-major (ru = r . G).  halfG: SplineR2R bc_sign parities (NULL = Gamma).  kcart: [n_orb][3]
 * Cartesian twist vectors for C2C (NULL for R2R).  finalizeConstruction() == this call (table pushed to HBM). */
int qmcb_spline_create(qmcb_spline** h, int precision, int kind, const int grid[3], int n_orb, int n_spl, size_t npad,
                       const void* coefs_host, const double G[9], const int halfG[3], const double* kcart);
int qmcb_spline_destroy(qmcb_spline* h);
size_t qmcb_spline_table_bytes(const qmcb_spline* h);
/* SPOSet::mw_evaluateValue (SPOSet.h:300).  r_host [nw][3] Cartesian doubles; psi_host [nw][n_orb] VT (complex interleaved for C2C) */
int qmcb_spline_mw_evaluate_value(qmcb_spline* h, int nw, const double* r_host, void* psi_host);
/* SPOSet::mw_evaluateVGL (SPOSet.h:313).  dpsi_host [nw][n_orb][3], d2psi_host [nw][n_orb] */
int qmcb_spline_mw_evaluate_vgl(qmcb_spline* h, int nw, const double* r_host, void* psi_host, void* dpsi_host,
                                void* d2psi_host);
/* SPOSet::mw_evaluateVGLandDetRatioGrads (SPOSet.h:346-352).  invrow_host [nw][ld_inv] VT.  Outputs:
 * phi_vgl_host [5][nw][n_orb] (may be NULL), ratios_host [nw], grads_host [nw][3] (already divided by ratio). */
int qmcb_spline_mw_evaluate_vgl_ratio_grads(qmcb_spline* h, int nw, const double* r_host, const void* invrow_host,
                                            size_t ld_inv, void* phi_vgl_host, void* ratios_host, void* grads_host);
/* SPOSet::mw_evaluateDetRatios (SPOSet.h:257): V-only evaluation at nvp virtual positions, each dotted with the
 * inverse row of its reference walker: ratios[i] = sum_j invrow[ref_walker[i]][j] * psi_j(r_vp[i]). */
int qmcb_spline_mw_evaluate_det_ratios(qmcb_spline* h, int nvp, const double* r_vp_host, const int* ref_walker_host,
                                       int n_ref, const void* invrow_host, size_t ld_inv, void* ratios_host);
/* device-resident variant used inside the sweep and by bench.py's kernel-only timing: positions and inverse rows
 * already in HBM, nothing copied, nothing synchronised.  r_dev [nw][3] RT; invrow_dev [nw][ld_inv]; phi_vgl_dev
 * [5][nw][n_orb]; rg_parts_dev [nw][qmcb_spline_rg_parts(h)][4] = partial sums of ratio, grad*ratio (x,y,z) --
 * undivided and to be added in index order, like the reference kernel's per-team partials which its host code adds
 * (SplineR2R.cpp:566-581).  stream: a cudaStream_t cast to void* (NULL = default stream). */
int qmcb_spline_mw_vgl_ratio_grads_dev(qmcb_spline* h, int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv,
                                       void* phi_vgl_dev, void* rg_parts_dev, void* stream);
int qmcb_spline_rg_parts(const qmcb_spline* h);
/* the contract of SPOSet::mw_evaluateVGLandDetRatioGrads for an offload SPOSet as DiracDeterminantBatched::mw_ratioGrad
 * drives it (DiracDeterminantBatched.cpp:334-346: isOMPoffload() == true): positions from the host (P_list[iw].activeR),
 * the inverse rows as DEVICE memory [nw][ld_inv] (mw_getInvRow(..., on_host = false)), phi_vgl_dev [5][nw][n_orb] = the
 * device copy of phi_vgl_v, current on return (may be NULL), ratios_host [nw] and grads_host [nw][3] VT on the host (grads
 * already divided by the ratio, SPOSet.cpp:171).  Synchronises `stream` (the crowd's queue) before returning.           */
int qmcb_spline_mw_evaluate_vgl_ratio_grads_offload(qmcb_spline* h, int nw, const double* r_host, const void* invrow_dev,
                                                    size_t ld_inv, void* phi_vgl_dev, void* ratios_host, void* grads_host,
                                                    void* stream);

/* ---- crowd: walker batch + trial wavefunction state ---------------------------------------------- */
typedef struct qmcb_system
{
  int precision;      /* QMCB_FULL / QMCB_MIXED                                                    */
  int n_up, n_dn;     /* electrons (= orbitals) per spin determinant                               */
  double lattice[9];  /* rows = cell vectors (bohr), periodic in all directions                     */
  qmcb_spline* spo[2];/* SPOSet of each spin determinant (may be the same handle)                   */
  int delay_rank;     /* slaterdeterminant delay_rank (Fermion/SlaterDetBuilder.cpp:402-409)        */
  /* two-body Jastrow, B-spline functors uu (= dd) and ud; n_j2 = 0 disables (Jastrow/TwoBodyJastrow.h:57) */
  int n_j2;
  const double* j2_uu;
  const double* j2_ud; /* NULL: the like-spin functor (cusp -1/4 included) serves every pair, as TwoBodyJastrow::addFunc does */
  double j2_rcut;
  /* one-body Jastrow (Jastrow/J1OrbitalSoA.h); nions = 0 disables */
  int nions;
  const double* ion_pos; /* [nions][3] */
  const int* ion_grp;    /* [nions]    */
  int n_ion_groups;
  int n_j1;
  const double* j1_params; /* [n_ion_groups][n_j1] */
  const double* j1_rcut;   /* [n_ion_groups]       */
} qmcb_system;

int qmcb_crowd_create(qmcb_crowd** c, const qmcb_system* sys, int nw);
int qmcb_crowd_destroy(qmcb_crowd* c);
int qmcb_crowd_sync(qmcb_crowd* c);
size_t qmcb_crowd_device_bytes(const qmcb_crowd* c);
/* 1 when the crowd was created over SplineC2C tables: the determinant value type VT is then complex (the reference's
 * QMC_COMPLEX build) and every VT / PsiValue / gradient buffer below holds interleaved (re, im) pairs:
 *   component level (qmcb_det_*): VT = complex<float> (mixed) or complex<double> (full);
 *   trial-wavefunction level (qmcb_twf_*): ratios [nw][2], grads [nw][3][2], G [nw][N][3][2], L [nw][N][2] doubles.
 * The Jastrow factors, positions, drifts and the kinetic energy stay real.                                          */
int qmcb_crowd_is_complex(const qmcb_crowd* c);
/* ParticleSet::R for every walker, [nw][N][3] doubles (loadWalker) */
int qmcb_crowd_set_positions(qmcb_crowd* c, const double* R_host);
int qmcb_crowd_get_positions(qmcb_crowd* c, double* R_host);

/* WaveFunctionComponent::mw_recompute / DiracDeterminantBatched::mw_recompute (DiracDeterminantBatched.cpp:1122-1198)
 * + TwoBodyJastrow::mw_recompute + J1: from-scratch psiM, FP64 inverse + log-determinant, Jastrow sums.          */
int qmcb_twf_mw_recompute(qmcb_crowd* c);
/* TrialWaveFunction::mw_evalGrad (TrialWaveFunction.cpp:568): grads_host [nw][3] doubles = sum over components.  */
int qmcb_twf_mw_eval_grad(qmcb_crowd* c, int iat, double* grads_host);
/* ParticleSet::mw_makeMove (Particle/ParticleSet.cpp:397): proposes R[iat] + displ for every walker and launches the
 * distance-table rows (SoaDistanceTableAAOMPTarget::mw_move :265-371); asynchronous.  displ_host [nw][3] doubles.  */
int qmcb_ps_mw_make_move(qmcb_crowd* c, int iat, const double* displ_host);
/* TrialWaveFunction::mw_calcRatioGrad (:685): ratios_host [nw] (PsiValue, double), grads_host [nw][3] doubles.    */
int qmcb_twf_mw_calc_ratio_grad(qmcb_crowd* c, int iat, double* ratios_host, double* grads_host);
/* TrialWaveFunction::mw_accept_rejectMove (:790) followed by ParticleSet::mw_accept_rejectMove (ParticleSet.cpp:717).
 * accepted_host [nw] (0/1).  safe_to_delay = 0 forces the Woodbury flush at once (DiracDeterminantBatched.cpp:519).   */
int qmcb_twf_mw_accept_reject(qmcb_crowd* c, int iat, const uint8_t* accepted_host, int safe_to_delay);
/* TrialWaveFunction::mw_completeUpdates (:833): flush pending delayed updates of both determinants.               */
int qmcb_twf_mw_complete_updates(qmcb_crowd* c);
/* TrialWaveFunction::mw_calcRatio (TrialWaveFunction.cpp:494-510; DiracDeterminantBatched::mw_calcRatio,
 * Fermion/DiracDeterminantBatched.cpp:742-789): the ratio of the move proposed by qmcb_ps_mw_make_move WITHOUT gradients
 * (sweeps without drift).  Only orbital values are gathered; an accept that follows leaves the determinant's gradient and
 * Laplacian rows stale (the reference's ORB_PBYP_RATIO mode) and they are re-evaluated from the committed positions when
 * next needed (qmcb_twf_mw_evaluate_gl, qmcb_twf_mw_eval_grad, a device sweep).  ratios_host [nw] doubles (x2 complex).     */
int qmcb_twf_mw_calc_ratio(qmcb_crowd* c, int iat, double* ratios_host);
/* TrialWaveFunction::mw_evaluateRatios (TrialWaveFunction.cpp:1079-1110) for the non-local pseudopotential: nvp virtual
 * positions; position i belongs to walker walker[i] and stands for its electron ref_ptcl[i] (VirtualParticleSet::refPtcl).
 * ratios_host[i] = product over the selected components of psi(electron at r_vp[i]) / psi: determinants through
 * DiracDeterminantBatched::mw_evaluateRatios (DiracDeterminantBatched.cpp:812-848; V-only spline gather dotted with the
 * row of psiMinv), Jastrows through TwoBodyJastrow::mw_evaluateRatios (Jastrow/TwoBodyJastrow.cpp:174-210,
 * BsplineFunctor::mw_evaluateV, Jastrow/BsplineFunctor.cpp:135-200) and J1OrbitalSoA::evaluateRatios.
 * compute_type: 0 ALL, 1 FERMIONIC, 2 NONFERMIONIC (TrialWaveFunction::ComputeType).  Pending delayed updates are applied
 * first (the reference evaluates the pseudopotential after mw_completeUpdates).  r_vp_host [nvp][3] Cartesian doubles;
 * ratios_host [nvp] doubles (x2 complex).                                                                              */
int qmcb_twf_mw_evaluate_ratios(qmcb_crowd* c, int nvp, const int* walker, const int* ref_ptcl, const double* r_vp_host,
                                int compute_type, double* ratios_host);
/* TrialWaveFunction::mw_evaluateGL (:869), fromscratch = false: G_host [nw][N][3], L_host [nw][N] (either may be
 * NULL), logpsi_host [nw] (real part of log psi), ke_host [nw] = -1/2 sum(L + G.G) (BareKineticEnergy).            */
int qmcb_twf_mw_evaluate_gl(qmcb_crowd* c, double* G_host, double* L_host, double* logpsi_host, double* ke_host);

/* ---- component level: DiracDeterminantBatched + DelayedUpdateBatched engine ----------------------- *
 * `spin` selects the determinant; `row` = iat - FirstIndex.                                           */
/* DelayedUpdateBatched::mw_evalGrad (Fermion/DelayedUpdateBatched.h:354-400): prepares the inverse rows
 * (mw_prepareInvRow :174-235) and returns grad_now = invRow . dpsiM[row]; grads_host [nw][3] VT.        */
int qmcb_det_mw_eval_grad(qmcb_crowd* c, int spin, int row, void* grads_host);
/* DelayedUpdateBatched::mw_getInvRow (:763-810): device pointer to the contiguous [nw][ld] current inverse rows
 * (prepared if needed); *ld receives the row stride.  invrow_host, if not NULL, receives a copy [nw][n].        */
int qmcb_det_mw_get_inv_row(qmcb_crowd* c, int spin, int row, const void** invrow_dev, size_t* ld, void* invrow_host);
/* DiracDeterminantBatched::mw_ratioGrad (DiracDeterminantBatched.cpp:308-354): spline VGL at the proposed positions
 * (set by qmcb_ps_mw_make_move) + ratio/grad against the current inverse rows.  ratios_host [nw] VT, grads_host [nw][3] VT */
int qmcb_det_mw_ratio_grad(qmcb_crowd* c, int spin, int row, void* ratios_host, void* grads_host);
/* DelayedUpdateBatched::mw_accept_rejectRow (:542-670) incl. pseudo-accept of rejected walkers and the flush when
 * delay_count reaches delay_rank (mw_updateInvMat :675-738).                                                        */
int qmcb_det_mw_accept_reject(qmcb_crowd* c, int spin, int row, const uint8_t* accepted_host);
/* DelayedUpdateBatched::mw_updateInvMat + mw_transferAinv_D2H: flush, then copy psiMinv [nw][n][n] VT (padding
 * stripped) and log-determinants [nw][2] (re, im) to the host; either output may be NULL.                            */
int qmcb_det_mw_complete_updates(qmcb_crowd* c, int spin, void* psiMinv_host, double* logdet_host);
/* test hook mirroring the unit tests with FakeSPO (test_DiracDeterminantBatched.cpp:262-470): load psiM/dpsiM/d2psiM
 * directly ([nw][n][n], [nw][n][n][3], [nw][n][n] VT) and invert.                                                      */
int qmcb_det_mw_recompute_from_matrices(qmcb_crowd* c, int spin, const void* psiM_host, const void* dpsiM_host,
                                        const void* d2psiM_host);
/* load externally computed orbital VGL of the proposed move ([5][nw][n] VT) in place of the spline evaluation
 * (FakeSPO-style determinant tests).                                                                                 */
int qmcb_det_set_phi_vgl(qmcb_crowd* c, int spin, const void* phi_vgl_host);
int qmcb_det_mw_ratio_grad_from_phi(qmcb_crowd* c, int spin, int row, void* ratios_host, void* grads_host);
int qmcb_det_delay_count(qmcb_crowd* c, int spin);
/* measurement hook for bench.py / scripts: runs `reps` back-to-back DelayedUpdateBatched::mw_updateInvMat launches
 * (Fermion/DelayedUpdateBatched.h:675-738) of determinant `spin` with `delay_count` pending slots and returns the mean
 * time of one flush in microseconds (CUDA events on the crowd's stream).  The slots hold whatever the last sweep left, so
 * the flushes are arithmetic on stale data; the inverse is saved before and restored afterwards, so the crowd is left as
 * qmcb_twf_mw_complete_updates leaves it.                                                                               */
int qmcb_det_time_update_inv_mat(qmcb_crowd* c, int spin, int delay_count, int reps, double* us_per_flush);
/* measurement hook: the FP64 inverse-transpose + log-determinant of the current Slater matrices of determinant `spin`
 * (DiracMatrixInverterCUDA::mw_invertTranspose, Fermion/DiracMatrixInverterCUDA.hpp:306-369), the step of
 * qmcb_twf_mw_recompute after the orbital evaluation; mean time of `reps` runs in microseconds, CUDA events on the crowd's
 * stream.  method 1: cublas<t>getrfBatched + getriBatched, the routines the reference calls
 * (detail/CUDA/cuBLAS_LU.cu:61-210); method 2: this library's blocked Gauss-Jordan kernels (csrc/inverse.cuh), the
 * default of qmcb_twf_mw_recompute (environment QMCB_LU=cublas selects the cuBLAS routines there).  The crowd is left
 * recomputed.                                                                                                            */
int qmcb_det_time_inverse(qmcb_crowd* c, int spin, int method, int reps, double* us_per_call);

/* ---- component level: distance rows + two-body Jastrow -------------------------------------------- */
/* SoaDistanceTableAAOMPTarget::mw_move temp/old rows after qmcb_ps_mw_make_move: [2][nw][4][N] RT (r,dx,dy,dz; new then old) */
int qmcb_dtaa_get_temp_rows(qmcb_crowd* c, void* rows_host);
/* TwoBodyJastrow::mw_ratioGrad (Jastrow/TwoBodyJastrow.cpp:542-578): ratios_host [nw] doubles, grads_host [nw][3] RT */
int qmcb_j2_mw_ratio_grad(qmcb_crowd* c, int iat, double* ratios_host, void* grads_host);
/* TwoBodyJastrow::mw_accept_rejectMove (:631-664) */
int qmcb_j2_mw_accept_reject(qmcb_crowd* c, int iat, const uint8_t* accepted_host);
/* per-particle sums of walker iw: Uat [N], dUat [3][N], d2Uat [N] as doubles */
int qmcb_j2_get_state(qmcb_crowd* c, int iw, double* Uat, double* dUat, double* d2Uat);

/* ---- device-resident driver: VMCBatched::advanceWalkers p-by-p loop (VMC/VMCBatched.cpp:106-176) --- *
 * One call = `nsteps` sweeps of all N electrons of all walkers with the accept/reject test done on the device from
 * a std::mt19937 stream (Utilities/StdRandom.h:34-48) and Box-Muller Gaussians (RandomSeqGenerator.h:33-52); no
 * host round trip inside the sweep.  accept_log_host (optional) [nsteps][N][nw].                                      */
typedef struct qmcb_vmc_params
{
  double tau;
  int use_drift;
  uint32_t seed;      /* the crowd's std::mt19937 seed */
  int use_cuda_graph; /* capture one sweep into a CUDA graph and replay it */
  /* 1: the move loop of DMCBatched::advanceWalkers (QMCDrivers/DMC/DMCBatched.cpp:142-262) instead of VMCBatched's:
   * a move is rejected when the ratio is zero or changes the phase (SFNBranch::phaseChanged, SFNBranch.h:161-169: real
   * wavefunctions reject a node crossing, complex ones never), prob = |ratio|^2 exp(log_gb - log_gf) is tested as a whole
   * against eps and the uniform, and tau |delta|^2 is accumulated per walker for proposed and accepted moves.          */
  int dmc;
  /* which kernels run the sweep.  0: automatic -- the persistent walker-segment kernel (one CTA per walker for the moves
   * between two Woodbury flushes: proposal, spline gather, ratio, Metropolis test, accept and next-row preparation in one
   * launch; csrc/segment.cuh) whenever the wavefunction is eligible (real orbitals, at most 384 per spin) and every
   * walker's CTA can be resident at once, otherwise the two-kernel path (boundary kernel + spline gather per move).
   * 1: always the two-kernel path.  2: the segment kernel or an error.  Both give the same acceptance sequence.           */
  int sweep_kernel;
} qmcb_vmc_params;
int qmcb_vmc_init(qmcb_crowd* c, const qmcb_vmc_params* p);
int qmcb_vmc_sweep(qmcb_crowd* c, int nsteps, uint8_t* accept_log_host);
/* asynchronous launch of one sweep on the crowd's stream (bench.py brackets it with CUDA events) */
int qmcb_vmc_sweep_async(qmcb_crowd* c);
/* accepted / rejected moves PER WALKER since qmcb_vmc_init: n_accept [nw], n_reject [nw] */
int qmcb_vmc_counts(qmcb_crowd* c, long long* n_accept, long long* n_reject);
/* 2 when the sweeps of this crowd run on the persistent walker-segment kernel, 1 on the two-kernel path, 0 before
 * qmcb_vmc_init                                                                                                          */
int qmcb_vmc_sweep_kernel(qmcb_crowd* c);
/* how the per-electron calls (qmcb_twf_mw_eval_grad / qmcb_ps_mw_make_move / qmcb_twf_mw_calc_ratio_grad /
 * qmcb_twf_mw_accept_reject) of this crowd are served: 2 = by a resident walker-segment kernel through host mailboxes
 * (one launch per segment of <= delay_rank moves; engaged when the calls arrive in the driver loop's order, real
 * orbitals, <= 384 per spin; environment QMCB_HOST_KERNEL=launch disables it), 1 = one or two launches per call,
 * 0 = no per-electron call has been made yet.  Both give the same results.                                              */
int qmcb_crowd_host_kernel(qmcb_crowd* c);
/* measurement hook: one sweep outside the CUDA graph with an event pair around every launch.  out9[0] = sweep time (us);
 * out9[1..4] = summed time of the walker-segment kernel, the boundary kernel, the spline gather and the Woodbury flush;
 * out9[5..8] = their launch counts.                                                                                     */
int qmcb_vmc_profile_sweep(qmcb_crowd* c, double* out9);
/* DMC: per-walker rr_accepted / rr_proposed of the LAST sweep (walker Properties R2ACCEPTED / R2PROPOSED,
 * DMCBatched.cpp:139-140,191-222), [nw] doubles each.                                                                 */
int qmcb_dmc_get_rr(qmcb_crowd* c, double* rr_accepted_host, double* rr_proposed_host);

/* ---- walker state for branching and load balancing (WalkerControl::branch, QMCDrivers/DMC/WalkerControl.cpp:151-240;
 * MCPopulation::fissionHighMultiplicityWalkers; swapWalkersSimple :312-500 ships the walker's buffers between ranks).
 * The packed state is device resident: positions, both inverse matrices with their gradient/Laplacian rows and
 * log-determinants, the Jastrow sums -- what the reference keeps in the walker's DataSet so that the receiver does not
 * recompute.  Pending delayed updates are flushed first.  dev_buf must hold qmcb_crowd_walker_bytes() bytes.           */
size_t qmcb_crowd_walker_bytes(const qmcb_crowd* c);
int qmcb_crowd_pack_walker(qmcb_crowd* c, int iw, void* dev_buf);
int qmcb_crowd_unpack_walker(qmcb_crowd* c, int iw, const void* dev_buf);
/* duplicate walker src over walker dst inside the crowd (a copy made by branching) */
int qmcb_crowd_copy_walker(qmcb_crowd* c, int src, int dst);
/* live walker count of the crowd (1 <= n <= capacity = the nw given at creation): branching kills and spawns walkers
 * (MCPopulation::killWalker / spawnWalker); walkers [0, n) take part in every mw_* call and in the sweep.  The walker
 * indices of pack/unpack/copy address the whole capacity so that a spawned walker can be filled before it goes live.  */
int qmcb_crowd_set_num_walkers(qmcb_crowd* c, int n);
int qmcb_crowd_num_walkers(const qmcb_crowd* c);
int qmcb_crowd_capacity(const qmcb_crowd* c);
/* the crowd's cudaStream_t as void* so callers can record events on the launching stream */
void* qmcb_crowd_stream(qmcb_crowd* c);

#ifdef __cplusplus
}
#endif
#endif /* QMCB_H */
