/*
 * dgr_oracle.c — CPU oracle of the differentiable Gaussian-splat rasterizer path (forward + backward).
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the library built from this file.  The product (dreamgaussian_b200/) never does.
 *
 * PARITY UNPINNED: the reference (dreamgaussian) calls an un-vendored, un-pinned third-party CUDA package for
 * this path (pip: diff_gaussian_rasterization, https://github.com/ashawkey/diff-gaussian-rasterization, no commit
 * pinned: /root/reference/readme.md:30-32) and ships no tests or golden vectors.  The
 * algorithm restated here is that package's published one; see dgr_oracle_impl.h for the
 * in-tree reference lines each part is anchored on.
 *
 * Built twice into one shared object: float32 (same arithmetic type as the CUDA path; used as the CPU baseline)
 * and float64 (the parity checker).  OpenMP over tiles / Gaussians.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/dgr_constants.h"

#define REAL float
#define SUF f32
#include "dgr_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF f64
#include "dgr_oracle_impl.h"
#undef REAL
#undef SUF

int dgr_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void dgr_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
