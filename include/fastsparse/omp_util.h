/* omp_util.h -- drop-in for omp_util.h:9-28.  The products no longer run on OpenMP
 * threads (they run on the GPU); these wrappers remain because callers size scratch
 * with them (parallel_bcsr_AA_mul_B's ytmp, test_sparse.c:99-107, bench_csr.c:95-103).
 * Unlike the reference this header does not pull in <cblas.h>. */
#ifndef OMP_UTIL_H
#define OMP_UTIL_H
#include <stdio.h>
#if defined(_OPENMP)
#include <omp.h>
static inline int nthreads(void) { return omp_get_num_threads(); }
static inline int thread_num(void) { return omp_get_thread_num(); }
static inline int thread_limit(void) { return omp_get_max_threads(); }
static inline void threads_init(void) { printf("Using OpenMP with up to %d threads.\n", thread_limit()); }
#else
static inline int nthreads(void) { return 1; }
static inline int thread_num(void) { return 0; }
static inline int thread_limit(void) { return 1; }
static inline void threads_init(void) {}
#endif
#endif /* OMP_UTIL_H */
