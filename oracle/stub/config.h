/* oracle/stub/config.h -- minimal stand-in for the cmake-generated config.h of the reference
 * (ref: src/config.h.cmake.in) so that a handful of reference headers compile without its build
 * system.  Test infrastructure only.  Offload pragmas vanish exactly as they do in a CPU build of the
 * reference (config.h.cmake.in:36-42). */
#ifndef QMC_ORACLE_STUB_CONFIG_H
#define QMC_ORACLE_STUB_CONFIG_H
#define PRAGMA_OFFLOAD(x)
#define PRAGMA_OMP_TASKLOOP(x) _Pragma("omp taskgroup")
#define OHMMS_DIM 3
#define OHMMS_INDEXTYPE int
#define OHMMS_PRECISION double
#define OHMMS_PRECISION_FULL double
#define QMC_SIMD_ALIGNMENT 64
#define HAVE_SINCOS 1
#define HAVE_POSIX_MEMALIGN 1
#ifdef __cplusplus
#define restrict __restrict__
#endif
#endif
