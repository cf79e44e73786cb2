/* =====================================================================================================
 * include/qmcb_driver.h -- host-side batched VMC driver ABOVE the C ABI of include/qmcb.h.
 *
 * A from-scratch C++ restatement of the reference's driver loop for this path, i.e. the caller of the hot path:
 *   VMCBatched::advanceWalkers  (src/QMCDrivers/VMC/VMCBatched.cpp:58-226; p-by-p loop :106-176)
 *   one host thread per crowd   (VMCBatched.cpp:348,405: ParallelExecutor over crowds; docs/methods.rst:143-163)
 * It only calls the qmcb_* entry points with HOST buffers (positions, gradients, ratios and accept flags cross
 * PCIe every move exactly like the reference's flex_* dispatch), owns the per-crowd std::mt19937 stream
 * (Utilities/StdRandom.h:34-48), the Box-Muller Gaussians (Particle/ParticleBase/RandomSeqGenerator.h:33-52), the UNR
 * drift (GreenFunctionModifiers/DriftModifierUNR.cpp:20-31) and the Metropolis test (VMCBatched.cpp:152-167).
 * bench.py times this loop for the end-to-end ("e2e") number.
 * ===================================================================================================== */
#ifndef QMCB_DRIVER_H
#define QMCB_DRIVER_H
#include "qmcb.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct qmcb_host_vmc qmcb_host_vmc;
const char* qmcb_host_vmc_last_error(void);

/* crowds: ncrowds crowd handles of equal precision; seeds: one std::mt19937 seed per crowd (different crowd counts give
 * different streams, docs/methods.rst:155-157).  precision: QMCB_FULL / QMCB_MIXED (RealType of the driver scalars). */
int qmcb_host_vmc_create(qmcb_host_vmc** d, qmcb_crowd** crowds, const int* nw_per_crowd, int ncrowds, int n_electrons,
                         int precision, const uint32_t* seeds, double tau, int use_drift);
int qmcb_host_vmc_destroy(qmcb_host_vmc* d);
/* runs nsteps sweeps (sub_steps = 1) on all crowds concurrently, one host thread per crowd; blocks until done.
 * accept_log_host (optional): [nsteps][N][nw_total], walkers ordered crowd by crowd. */
int qmcb_host_vmc_run(qmcb_host_vmc* d, int nsteps, uint8_t* accept_log_host);
/* totals since creation */
int qmcb_host_vmc_counts(qmcb_host_vmc* d, long long* n_accept, long long* n_reject);
/* bytes moved per sweep through the C ABI by this driver: host->device and device->host (all crowds) */
int qmcb_host_vmc_bytes_per_sweep(qmcb_host_vmc* d, long long* h2d, long long* d2h);

/* =====================================================================================================
 * DMC layer above the device-resident sweep, C++ (csrc/dmc_host.cpp):
 *   DMCBatched::advanceWalkers, the part after the move loop      QMCDrivers/DMC/DMCBatched.cpp:264-292
 *   SFNBranch::branchWeight / setBranchCutoff / updateParamAfterPopControl (warm-up AND main stage, UNLIMITED_HISTORY
 *   reference energy)                                              QMCDrivers/SFNBranch.h:199-208, SFNBranch.cpp:133-215,290-320
 *   WalkerControl::branch (dynamic population), computeCurData, determineNewWalkerPopulation, swapWalkersSimple
 *                                                                  QMCDrivers/DMC/WalkerControl.cpp:151-313,312-500
 *   MCPopulation::killWalker / spawnWalker / fissionHighMultiplicityWalkers
 * The move loop itself (phase rejection, rr accumulators) runs on the device: qmcb_vmc_init(dmc = 1) + qmcb_vmc_sweep.
 * A walker's state never visits the host: copies inside a rank are device-to-device, walkers that change rank travel as
 * packed device buffers through the communicator the caller supplies.
 * The local energy of this layer is whatever the engine reports per walker (the harness: kinetic energy; the
 * Hamiltonian proper is out of scope); everything downstream of it is the reference's arithmetic.
 * ===================================================================================================== */

/* Communicator: what WalkerControl needs from `Communicate* myComm` (allreduce of curData; send_value + send_n /
 * receive_n of one walker buffer, WalkerControl.cpp:276-280,373,395-471).  The caller owns the transport -- NCCL through
 * torch.distributed in bench.py, MPI inside QMCPACK -- and two transfer buffers of qmcb_crowd_walker_bytes() bytes in the
 * memory space of the walkers (device memory for a qmcb_crowd).  send ships `send_buf` + the 4-double header, recv fills
 * `recv_buf` + header; both block; every rank walks the same schedule, so calls pair up.                               */
typedef struct qmcb_comm
{
  int rank, size;
  void* ctx;
  int (*allreduce_sum)(void* ctx, double* host_buf, int n);
  int (*send)(void* ctx, int dst, const double* header4);
  int (*recv)(void* ctx, int src, double* header4);
  void* send_buf;
  void* recv_buf;
} qmcb_comm;

/* The walker batch the DMC layer drives.  qmcb_dmc_create() fills it for a qmcb_crowd; tests supply their own. */
typedef struct qmcb_dmc_engine
{
  void* ctx;
  int (*num_walkers)(void* ctx);
  int (*capacity)(void* ctx);
  int (*sweep)(void* ctx);                                        /* one DMC move loop over all electrons            */
  int (*local_energies)(void* ctx, double* e);                    /* [nw]                                            */
  int (*get_rr)(void* ctx, double* rr_accepted, double* rr_proposed); /* [nw] each, of the last sweep                */
  int (*copy_walker)(void* ctx, int src, int dst);
  int (*set_num_walkers)(void* ctx, int n);
  int (*pack_walker)(void* ctx, int iw, void* buf);
  int (*unpack_walker)(void* ctx, int iw, const void* buf);
} qmcb_dmc_engine;

typedef struct qmcb_dmc_params
{
  double tau;
  int target_walkers;         /* global target population (SFNBranch iParam[B_TARGETWALKERS]); 0: the initial one      */
  uint32_t branch_seed;       /* WalkerControl's own std::mt19937                                                      */
  double sigma2;              /* initial variance estimate (vParam[SIGMA2]); 10 in SFNBranch's constructor             */
  double sigma_bound;         /* vParam[SIGMA_BOUND] = 10                                                              */
  double feedback;            /* vParam[FEEDBACK]                                                                      */
  int warmup_steps;           /* iParam[B_WARMUPSTEPS]                                                                 */
  int energy_update_interval; /* iParam[B_ENERGYUPDATEINTERVAL]                                                        */
  int use_tau_eff;            /* BranchMode[B_USETAUEFF]                                                               */
} qmcb_dmc_params;

typedef struct qmcb_dmc_ensemble
{
  /* MCDataType ensemble_property_ of this generation (WalkerControl.cpp:104-117) */
  double energy, variance, weight, num_samples, r2_accepted, r2_proposed, living_fraction;
  /* branch engine after updateParamAfterPopControl */
  double e_trial, e_ref, tau_eff, branch_cutoff;
  int population; /* global, after branching */
  int local;      /* this rank, after branching and load balancing */
  long long walkers_sent, walkers_received, bytes_sent, bytes_received; /* this rank, this generation */
} qmcb_dmc_ensemble;

typedef struct qmcb_dmc qmcb_dmc;
const char* qmcb_dmc_last_error(void);
/* crowd: initialised with qmcb_vmc_init(dmc = 1).  comm: NULL for a single rank. */
int qmcb_dmc_create(qmcb_dmc** d, qmcb_crowd* crowd, const qmcb_dmc_params* p, const qmcb_comm* comm);
int qmcb_dmc_create_with_engine(qmcb_dmc** d, const qmcb_dmc_engine* engine, const qmcb_dmc_params* p, const qmcb_comm* comm);
int qmcb_dmc_destroy(qmcb_dmc* d);
/* one generation: DMCBatched::advanceWalkers + WalkerControl::branch(iter, pop, iter == 0) +
 * SFNBranch::updateParamAfterPopControl                                                                                */
int qmcb_dmc_step(qmcb_dmc* d, int iter, qmcb_dmc_ensemble* out);
/* the two halves separately (tests) */
int qmcb_dmc_advance(qmcb_dmc* d);
int qmcb_dmc_branch(qmcb_dmc* d, int iter, int do_not_branch, qmcb_dmc_ensemble* out);
/* per-walker host state of the live walkers: Weight, LOCALENERGY, Age; returns the live count */
int qmcb_dmc_get_walkers(qmcb_dmc* d, double* weights, double* energies, long long* ages, int max_n);
/* overwrite Weight of the live walkers (tests) */
int qmcb_dmc_set_weights(qmcb_dmc* d, const double* weights, int n);
/* SFNBranch::branchWeight(enew, eold) with the current parameters */
double qmcb_dmc_branch_weight(qmcb_dmc* d, double enew, double eold);

#ifdef __cplusplus
}
#endif
#endif
