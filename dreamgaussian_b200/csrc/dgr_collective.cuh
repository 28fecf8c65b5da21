// dgr_collective.cuh — all-reduce(sum) of the flat per-Gaussian gradient buffer over NVLink, written for this path:
// the views of one iteration are sharded across GPUs and their gradients meet in ONE float32 buffer that
// lives in symmetric (peer-mapped) memory.  Two variants, both two-shot (rank r reduces slice r, then publishes it):
//   * multimem:  multimem.ld_reduce pulls the sum of all ranks' copies out of the NVSwitch (in-switch reduction, NVLS)
//                and multimem.st broadcasts the result — 2 x slice bytes per GPU cross NVLink;
//   * p2p:       plain peer loads of the slice from every rank (fixed order -> bit-identical result on every rank) and
//                peer stores of the sum to every rank.
// Every thread keeps kUnroll independent 16-byte operations per peer in flight (NVLink round trips are ~2-4 us).
// The two cross-rank barriers an all-reduce needs (nobody reads a peer's buffer before that peer has finished writing
// it; nobody leaves before every peer has stored its slice everywhere) are part of the kernel: block b of every rank
// signals block b of every peer through a small flag array in the same symmetric allocation (release store of a
// monotonically increasing epoch, acquire spin) — no extra launches, no host involvement.
#pragma once
#include "dgr_common.cuh"

namespace dgr {

constexpr int kMaxPeers = 16;
constexpr int kMaxFlagBlocks = 1024;                    // grid limit of the barrier-carrying kernels
constexpr int kUnroll = 4;
struct PeerPtrs { float *p[kMaxPeers]; };
struct PeerFlags { unsigned *p[kMaxPeers]; };          // per rank: u32 [2 phases][kMaxFlagBlocks][kMaxPeers]

__device__ __forceinline__ float4 ld_sys_f4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float *p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Block-level cross-rank barrier: thread w < world signals peer w and waits for peer w's signal (same block index there).
// Everything this block wrote before is visible system-wide before the signal (fence + release); everything the peers wrote
// before THEIR signal is visible to this block after the wait (acquire + block barrier).
__device__ __forceinline__ void peer_barrier(const PeerFlags &flags, int world, int rank, unsigned epoch, int phase) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        const size_t slot = ((size_t)phase * kMaxFlagBlocks + blockIdx.x) * kMaxPeers;
        st_release_sys_u32(flags.p[threadIdx.x] + slot + rank, epoch);
        const unsigned *mine = flags.p[rank] + slot + threadIdx.x;
        while ((int)(ld_acquire_sys_u32(mine) - epoch) < 0) { }
    }
    __syncthreads();
}

template <bool BAR>
__global__ void __launch_bounds__(512)
allreduce_p2p_kernel(PeerPtrs peers, PeerFlags flags, unsigned epoch, int world, int rank, size_t n4) {
    if (BAR) peer_barrier(flags, world, rank, epoch, 0);
    const size_t per = (n4 + world - 1) / world;
    const size_t s = (size_t)rank * per, e = (s + per < n4) ? (s + per) : n4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = s + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < e; i0 += kUnroll * stride) {
        float4 acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = i0 + u * stride;
            acc[u] = (i < e) ? ld_sys_f4(peers.p[0] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int w = 1; w < world; w++) {
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = i0 + u * stride;
                v[u] = (i < e) ? ld_sys_f4(peers.p[w] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
        for (int w = 0; w < world; w++) {
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = i0 + u * stride;
                if (i < e) st_sys_f4(peers.p[w] + 4 * i, acc[u]);
            }
        }
    }
    if (BAR) peer_barrier(flags, world, rank, epoch, 1);
}

template <bool BAR>
__global__ void __launch_bounds__(512)
allreduce_multimem_kernel(float *mc, PeerFlags flags, unsigned epoch, int world, int rank, size_t n4) {
    if (BAR) peer_barrier(flags, world, rank, epoch, 0);
    const size_t per = (n4 + world - 1) / world;
    const size_t s = (size_t)rank * per, e = (s + per < n4) ? (s + per) : n4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = s + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < e; i0 += kUnroll * stride) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = i0 + u * stride;
            if (i < e) {
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * i) : "memory");
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = i0 + u * stride;
            if (i < e)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
        }
    }
    if (BAR) peer_barrier(flags, world, rank, epoch, 1);
}

}  // namespace dgr
