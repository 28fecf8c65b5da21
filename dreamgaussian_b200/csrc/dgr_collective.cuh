// dgr_collective.cuh — all-reduce(sum) of the flat per-Gaussian gradient buffer over NVLink, written for this path:
// the views of one iteration are sharded across GPUs and their gradients meet in ONE float32 buffer that
// lives in symmetric (peer-mapped) memory.  Two variants, both two-shot (rank r reduces slice r, then publishes it):
//   * multimem:  multimem.ld_reduce pulls the sum of all ranks' copies out of the NVSwitch (in-switch reduction, NVLS)
//                and multimem.st broadcasts the result — 2 x slice bytes per GPU cross NVLink;
//   * p2p:       plain peer loads of the slice from every rank (fixed order -> bit-identical result on every rank) and
//                peer stores of the sum to every rank.
// Every thread keeps kUnroll independent 16-byte operations per peer in flight (NVLink round trips are ~2-4 us).
// The two cross-rank barriers an all-reduce needs (nobody reads a peer's buffer before that peer has finished writing
// it; nobody leaves before every peer has stored its slice everywhere) can be part of the kernel (peer_begin / peer_end
// below: flag words in the same symmetric allocation, no extra launches, no host involvement).
#pragma once
#include "dgr_common.cuh"

namespace dgr {

constexpr int kMaxFlagBlocks = 1024;                    // grid limit of the barrier-carrying kernels
constexpr int kUnroll = 4;
struct PeerPtrs { float *p[kMaxPeers]; };
struct PeerFlags { unsigned *p[kMaxPeers]; };          // per rank: a zero-initialised flag area of dgr_peer_flag_bytes()

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

// The two cross-rank barriers, at ONE remote store per peer each (flag words in the same symmetric allocation, a monotonically
// increasing epoch per call):
//   * begin: this kernel follows the rank's backward in stream order, so the rank's own buffer is final when the kernel starts;
//     block 0 tells every peer so, and every block polls the `world` words the peers write HERE (local memory) before it touches
//     a peer's buffer;
//   * end:   every block fences its peer stores system-wide and counts itself off; the last block tells every peer "my slice is
//     in your buffer" and stays until every peer has said the same — the kernel's completion then means the whole sum is
//     here, and all other blocks have long left the SMs.
// (A first version paired block b of every rank with block b of every peer, 2 x grid remote flag stores and system fences per
// rank: 20 us slower than host-launched barrier kernels at 2 GPUs, profiles/r2_earlier/r2_bench_n2_inkernel.json.)
constexpr int kSigBegin = 0, kSigEnd = kMaxPeers, kSigCount = 2 * kMaxPeers;      // word offsets inside a rank's flag area

// (Polling with relaxed loads and one fence after the word has arrived, signalling with relaxed stores behind an explicit fence,
// measured 9 us SLOWER than the acquire / release forms below at 2 GPUs — profiles/r2_phases_n2_*.json.)
// A peer that never arrives (its process died, the ranks disagree on the number of calls) must not hang the GPU: after
// kPeerWaitNs of polling the kernel traps, which surfaces as a CUDA error on the host instead of a stuck device.
constexpr unsigned long long kPeerWaitNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void wait_epoch(const unsigned *word, unsigned epoch) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while ((int)(ld_acquire_sys_u32(word) - epoch) < 0) {
        if ((++spins & 0x3fffu) == 0) {
            const unsigned long long t = global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > kPeerWaitNs) __trap();
        }
    }
}

__device__ __forceinline__ void peer_begin(const PeerFlags &flags, int world, int rank, unsigned epoch) {
    if ((int)threadIdx.x < world) {
        if (blockIdx.x == 0) st_release_sys_u32(flags.p[threadIdx.x] + kSigBegin + rank, epoch);
        wait_epoch(flags.p[rank] + kSigBegin + threadIdx.x, epoch);
    }
    __syncthreads();
}

__device__ __forceinline__ void peer_end(const PeerFlags &flags, int world, int rank, unsigned epoch) {
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                   // this block's peer stores are performed system-wide
        unsigned *cnt = flags.p[rank] + kSigCount;
        const unsigned old = atomicAdd(cnt, 1u);
        last = old == gridDim.x - 1;
        if (last) { __threadfence(); *cnt = 0u; }
    }
    __syncthreads();
    if (last && (int)threadIdx.x < world) {
        st_release_sys_u32(flags.p[threadIdx.x] + kSigEnd + rank, epoch);
        wait_epoch(flags.p[rank] + kSigEnd + threadIdx.x, epoch);
    }
}

template <bool BAR>
__global__ void __launch_bounds__(512)
allreduce_p2p_kernel(PeerPtrs peers, PeerFlags flags, unsigned epoch, int world, int rank, size_t n4) {
    if (BAR) peer_begin(flags, world, rank, epoch);
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
    if (BAR) peer_end(flags, world, rank, epoch);
}

// In-switch variant.  The two halves load the links in OPPOSITE directions: multimem.ld_reduce makes the switch fetch the slice from
// every GPU (each GPU sends its whole buffer out, receives one reduced slice), multimem.st sends one slice out and brings every
// rank's slice in.  Issuing all loads and then all stores leaves each direction idle half of the time, so a thread's kUnroll
// pieces are software-pipelined: the load of piece u+1 is in flight before piece u is stored.
template <bool BAR, bool PIPE>
__global__ void __launch_bounds__(512)
allreduce_multimem_kernel(float *mc, PeerFlags flags, unsigned epoch, int world, int rank, size_t n4) {
    if (BAR) peer_begin(flags, world, rank, epoch);
    const size_t per = (n4 + world - 1) / world;
    const size_t s = (size_t)rank * per, e = (s + per < n4) ? (s + per) : n4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = s + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < e; i0 += kUnroll * stride) {
        float4 v[kUnroll];
        if (!PIPE) {                                   // all loads, then all stores (A/B reference: DGR_AR_PIPELINE=0)
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = i0 + u * stride;
                if (i < e)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * i) : "memory");
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = i0 + u * stride;
                if (i < e)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
            }
            continue;
        }
#pragma unroll
        for (int u = 0; u <= kUnroll; u++) {
            if (u < kUnroll) {
                const size_t i = i0 + u * stride;
                if (i < e)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * i) : "memory");
            }
            if (u > 0) {
                const size_t i = i0 + (u - 1) * stride;
                if (i < e)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v[u - 1].x), "f"(v[u - 1].y), "f"(v[u - 1].z), "f"(v[u - 1].w) : "memory");
            }
        }
    }
    if (BAR) peer_end(flags, world, rank, epoch);
}

// ---- reduce-scatter by pushing, fused into the per-Gaussian backward (opt-in: DGR_PUSH=1) -----------------------------------------
// MEASURED NEGATIVE at 2 GPUs (profiles/r2_phases_n2_push.json): the per-Gaussian backward is ONE wave of CTAs whose stores all
// come at the end of the thread, so the pushed 12.5 MB do not overlap its arithmetic — the kernel grows by 29 us while this
// reduce/publish kernel is only 15-20 us shorter than the full two-shot all-reduce.  Kept as a tested alternative.
// Gaussian g belongs to rank g / per.  The backward of a rank's LAST view of the iteration stores the gradient rows of Gaussians it
// does not own into its slot of the owner's staging area (PeerPush, dgr_backward.cuh): the reduce-scatter traffic crosses NVLink
// while that kernel computes, and what is left for the collective is this kernel — the owner adds the `world` staged copies of its
// rows (local HBM reads, fixed rank order: every rank ends up with the same bits) and publishes the sums to every rank's buffer
// (multimem.st through the NVSwitch, or peer stores).  The flat buffer is a handful of [P, stride] segments, so a Gaussian range
// is one float range per segment.
struct FlatSegs { int n; long long off[8]; int stride[8]; };

template <int VEC>      // 4: the range starts on a 16-byte boundary and is a multiple of 4 floats long; 1: any range
__device__ __forceinline__ void reduce_publish_range(const PeerPtrs &peers, float *mc, const float *stage, size_t padded, int world, int rank,
                                                     size_t f0, size_t n) {
    const size_t nv = n / VEC, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nv; i0 += kUnroll * stride) {
        float4 acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < world; s++) {
            const float *src = (s == rank) ? peers.p[rank] : stage + (size_t)s * padded;        // own rows never left the local buffer
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = i0 + u * stride;
                if (i < nv) {
                    if (VEC == 4) { const float4 v = *reinterpret_cast<const float4 *>(src + f0 + 4 * i); acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w; }
                    else acc[u].x += src[f0 + i];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = i0 + u * stride;
            if (i >= nv) continue;
            if (VEC == 4) {
                if (mc) asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + f0 + 4 * i), "f"(acc[u].x), "f"(acc[u].y), "f"(acc[u].z), "f"(acc[u].w) : "memory");
                else for (int w = 0; w < world; w++) st_sys_f4(peers.p[w] + f0 + 4 * i, acc[u]);
            } else {
                if (mc) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + f0 + i), "f"(acc[u].x) : "memory");
                else for (int w = 0; w < world; w++) asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(peers.p[w] + f0 + i), "f"(acc[u].x) : "memory");
            }
        }
    }
}

template <bool BAR>
__global__ void __launch_bounds__(512)
reduce_staged_kernel(PeerPtrs peers, float *mc, const float *stage, size_t padded, PeerFlags flags, unsigned epoch, int world, int rank,
                     int P, int per, FlatSegs segs) {
    if (BAR) peer_begin(flags, world, rank, epoch);              // every rank's pushes into this rank's staging area have landed
    const long long g0 = (long long)rank * per, g1 = min((long long)P, g0 + per);
    if (g1 > g0) {
        for (int k = 0; k < segs.n; k++) {
            const size_t f0 = (size_t)(segs.off[k] + g0 * segs.stride[k]), n = (size_t)((g1 - g0) * segs.stride[k]);
            if (((f0 | n) & 3) == 0) reduce_publish_range<4>(peers, mc, stage, padded, world, rank, f0, n);
            else reduce_publish_range<1>(peers, mc, stage, padded, world, rank, f0, n);
        }
    }
    if (BAR) peer_end(flags, world, rank, epoch);
}

// The same push without a backward to ride on (a rank that rendered no view this iteration, or a buffer filled by something else):
// copies the rows this rank does not own from its local buffer into its slot at their owners.
__global__ void __launch_bounds__(512)
push_flat_kernel(const float *local, PeerPush push, int world, int rank, int P, FlatSegs segs) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int o = 0; o < world; o++) {
        if (o == rank) continue;
        const long long g0 = (long long)o * push.per, g1 = min((long long)P, g0 + push.per);
        if (g1 <= g0) continue;
        float *dst = const_cast<float *>(local) + push.delta[o];
        for (int k = 0; k < segs.n; k++) {
            const size_t f0 = (size_t)(segs.off[k] + g0 * segs.stride[k]), n = (size_t)((g1 - g0) * segs.stride[k]);
            if (((f0 | n) & 3) == 0) {
                for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride)
                    st_sys_f4(dst + f0 + 4 * i, *reinterpret_cast<const float4 *>(local + f0 + 4 * i));
            } else {
                for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
                    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + f0 + i), "f"(local[f0 + i]) : "memory");
            }
        }
    }
}

}  // namespace dgr
