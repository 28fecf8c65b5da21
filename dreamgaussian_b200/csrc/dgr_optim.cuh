// dgr_optim.cuh — SURVEY.md §8 row f2: the optimiser step of the stage-1 loop.  The reference builds
// torch.optim.Adam(l, lr=0.0, eps=1e-15) over six single-tensor parameter groups (/root/reference/gs_renderer.py:361-370)
// and steps it once per iteration (main.py:274-276): six groups x ~6 element-wise kernels each.  Here: ONE launch over
// the virtual concatenation of all groups, each with its own learning rate; float4 where alignment allows.
//   m <- m + (1 - b1) (g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (torch.optim.Adam, no weight decay, no amsgrad; bias corrections computed on the host in double).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dgr {

constexpr int kAdamMaxGroups = 8;
struct AdamGroups {
    float *param[kAdamMaxGroups];
    const float *grad[kAdamMaxGroups];
    float *m[kAdamMaxGroups];
    float *v[kAdamMaxGroups];
    unsigned long long start[kAdamMaxGroups + 1];     // prefix of element counts, in units of 4 elements (rounded up per group)
    unsigned long long n[kAdamMaxGroups];
    float step_size[kAdamMaxGroups];                  // lr / (1 - b1^t), t = that tensor's own step count (as torch keeps it)
    float inv_sqrt_bc2[kAdamMaxGroups];               // 1 / sqrt(1 - b2^t)
    int n_groups;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamGroups G, float one_minus_beta1, float beta2, float one_minus_beta2, float eps) {
    const unsigned long long total = G.start[G.n_groups];
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (unsigned long long)gridDim.x * blockDim.x) {
        int g = 0;
#pragma unroll
        for (int k = 1; k < kAdamMaxGroups; k++) g += (k < G.n_groups && q >= G.start[k]) ? 1 : 0;
        const unsigned long long e0 = (q - G.start[g]) * 4;
        const unsigned long long n = G.n[g];
        float *p = G.param[g] + e0, *m = G.m[g] + e0, *v = G.v[g] + e0;
        const float *gr = G.grad[g] + e0;
        const float ss = G.step_size[g], inv_sqrt_bc2 = G.inv_sqrt_bc2[g];
        const int cnt = (int)(n - e0 < 4 ? n - e0 : 4);
        const bool vec = cnt == 4 && ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)gr) & 15) == 0);
        float pv[4], mv[4], vv[4], gv[4];
        if (vec) {
            const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(m),
                         c = *reinterpret_cast<const float4 *>(v), d = *reinterpret_cast<const float4 *>(gr);
            pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; mv[0] = b.x; mv[1] = b.y; mv[2] = b.z; mv[3] = b.w;
            vv[0] = c.x; vv[1] = c.y; vv[2] = c.z; vv[3] = c.w; gv[0] = d.x; gv[1] = d.y; gv[2] = d.z; gv[3] = d.w;
        } else {
            for (int k = 0; k < cnt; k++) { pv[k] = p[k]; mv[k] = m[k]; vv[k] = v[k]; gv[k] = gr[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < cnt) {
                mv[k] = mv[k] + one_minus_beta1 * (gv[k] - mv[k]);          // 1 - beta computed in double on the host, as python does
                vv[k] = beta2 * vv[k] + one_minus_beta2 * gv[k] * gv[k];
                pv[k] = pv[k] - ss * (mv[k] / (sqrtf(vv[k]) * inv_sqrt_bc2 + eps));
            }
        }
        if (vec) {
            *reinterpret_cast<float4 *>(p) = make_float4(pv[0], pv[1], pv[2], pv[3]);
            *reinterpret_cast<float4 *>(m) = make_float4(mv[0], mv[1], mv[2], mv[3]);
            *reinterpret_cast<float4 *>(v) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        } else {
            for (int k = 0; k < cnt; k++) { p[k] = pv[k]; m[k] = mv[k]; v[k] = vv[k]; }
        }
    }
}

}  // namespace dgr
