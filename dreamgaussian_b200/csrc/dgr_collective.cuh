// dgr_collective.cuh — all-reduce(sum) of the flat per-Gaussian gradient buffer over NVLink, written for this path:
// the views of one iteration are sharded across GPUs and their gradients meet in ONE float32 buffer that
// lives in symmetric (peer-mapped) memory.  Two variants, both two-shot (rank r reduces slice r, then publishes it):
//   * multimem:  multimem.ld_reduce pulls the sum of all ranks' copies out of the NVSwitch (in-switch reduction, NVLS)
//                and multimem.st broadcasts the result — 2 x slice bytes per GPU cross NVLink;
//   * p2p:       plain peer loads of the slice from every rank (fixed order -> bit-identical result on every rank) and
//                peer stores of the sum to every rank.
// The caller brackets the kernel with two cross-rank barriers (symmetric-memory signal pads).
#pragma once
#include "dgr_common.cuh"

namespace dgr {

constexpr int kMaxPeers = 16;
struct PeerPtrs { float *p[kMaxPeers]; };

__device__ __forceinline__ float4 ld_sys_f4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float *p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(512)
allreduce_p2p_kernel(PeerPtrs peers, int world, int rank, size_t n4) {
    const size_t per = (n4 + world - 1) / world;
    const size_t s = (size_t)rank * per, e = (s + per < n4) ? (s + per) : n4;
    for (size_t i = s + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = ld_sys_f4(peers.p[0] + 4 * i);
        for (int w = 1; w < world; w++) {
            const float4 v = ld_sys_f4(peers.p[w] + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int w = 0; w < world; w++) st_sys_f4(peers.p[w] + 4 * i, acc);
    }
}

__global__ void __launch_bounds__(512)
allreduce_multimem_kernel(float *mc, int world, int rank, size_t n4) {
    const size_t per = (n4 + world - 1) / world;
    const size_t s = (size_t)rank * per, e = (s + per < n4) ? (s + per) : n4;
    for (size_t i = s + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        float *a = mc + 4 * i;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

}  // namespace dgr
