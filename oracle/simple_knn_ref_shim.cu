// Test infrastructure only (see oracle/README.md): a C entry point around the REFERENCE's own 3-NN code, compiled from
// the sources where they lie (/root/reference/simple-knn/simple_knn.cu, unmodified, by oracle/Makefile `ref`) into
// oracle/_ref/libsimple_knn_ref.so.  Used by tests/ and tools/ as the checker / baseline of row f3 — never by the product.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_dist_cuda2(int P, const float *d_points, float *d_mean_dists) {
    // spatial.cu:15-26 `distCUDA2`: means = full({P}, 0); SimpleKNN::knn(P, points, means)
    cudaMemset(d_mean_dists, 0, sizeof(float) * (size_t)P);
    SimpleKNN::knn(P, (float3 *)d_points, d_mean_dists);
    return (int)cudaDeviceSynchronize();
}
