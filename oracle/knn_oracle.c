/* knn_oracle.c — CPU restatement of the reference's distCUDA2.  TEST INFRASTRUCTURE ONLY.
 *
 * Reference: /root/reference/simple-knn/spatial.cu:15-26 (distCUDA2) -> simple_knn.cu:185-221 (SimpleKNN::knn) ->
 * simple_knn.cu:132-183 (boxMeanDist): for every point the three smallest squared distances to the OTHER points
 * (the point itself is skipped by index, :149/:171, so coincident points do count with distance 0), kept in ascending
 * order by updateKBest<3> (:118-130), and the result (best[0] + best[1] + best[2]) / 3.0f (:182).  The Morton sort
 * (:39-68, :203-212) and the 1024-point boxes with their min/max rejection test (:75-116, :153-176) only prune an
 * exact search — they never change the three winners — so the restatement is a brute-force scan.  Distances in float32
 * like the reference (dx*dx + dy*dy + dz*dz; nvcc may contract this into FMAs, which is why the parity tolerance of
 * the float32 result is 2 ulp-ish, 1e-6 relative, and the float64 variant below is the tie-breaker).
 *
 * PARITY PINNED for this row: tests/test_knn_gpu.py runs the reference's own simple_knn.cu (compiled unmodified into
 * oracle/_ref/libsimple_knn_ref.so by `make -C oracle ref`) next to this file and to the product kernel.
 */
#include <float.h>
#include <stddef.h>

static inline void update3f(float d, float *best) {            /* simple_knn.cu:118-130 */
    for (int j = 0; j < 3; j++)
        if (best[j] > d) { float t = best[j]; best[j] = d; d = t; }
}
static inline void update3d(double d, double *best) {
    for (int j = 0; j < 3; j++)
        if (best[j] > d) { double t = best[j]; best[j] = d; d = t; }
}

/* float32 arithmetic, as the reference */
void knn_oracle_f32(int P, const float *pts, float *mean_dists) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < P; i++) {
        float best[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
        const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        for (int j = 0; j < P; j++) {
            if (j == i) continue;
            const float dx = pts[3 * (size_t)j] - x, dy = pts[3 * (size_t)j + 1] - y, dz = pts[3 * (size_t)j + 2] - z;
            update3f(dx * dx + dy * dy + dz * dz, best);
        }
        mean_dists[i] = (best[0] + best[1] + best[2]) / 3.0f;      /* simple_knn.cu:182 */
    }
}

/* float64 evaluation of the same definition (float32 inputs): the yardstick both float32 implementations are held to */
void knn_oracle_f64(int P, const float *pts, double *mean_dists) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < P; i++) {
        double best[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        for (int j = 0; j < P; j++) {
            if (j == i) continue;
            const double dx = pts[3 * (size_t)j] - x, dy = pts[3 * (size_t)j + 1] - y, dz = pts[3 * (size_t)j + 2] - z;
            update3d(dx * dx + dy * dy + dz * dz, best);
        }
        mean_dists[i] = (best[0] + best[1] + best[2]) / 3.0;
    }
}
