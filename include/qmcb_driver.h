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

#ifdef __cplusplus
}
#endif
#endif
