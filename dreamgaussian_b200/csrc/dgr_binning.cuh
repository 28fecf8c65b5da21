// dgr_binning.cuh — tile binning (A3 of SURVEY.md §8a) without a global sort and without global atomics.
//
// The reference op sorts ALL (tile, depth) keys with one device-wide 64-bit radix sort (6 passes over N_inst pairs)
// and reads the instance count back to the host.  Here:
//   1. the preprocess kernel histograms the instances of each block of Gaussians per tile in shared memory and writes
//      one row of a [blocks x tiles] count matrix;
//   2. tile_colscan_kernel turns every column into exclusive per-(block, tile) offsets and the per-tile totals,
//      tile_scan_kernel turns the totals into per-tile [start, end) ranges (clipped to the buffer capacity) and
//      publishes the instance count — all on the device, no host round trip is needed to continue;
//   3. emit_instances_kernel (same block <-> Gaussian mapping) loads its matrix row into shared memory and appends
//      (depth bits << 32 | Gaussian id) keys at range.start + offset[block][tile] + local rank (shared-memory atomics);
//   4. tile_sort_gather_kernel sorts each tile's segment by (depth, id) in shared memory (bitonic network; keys are
//      unique, so the result is deterministic and identical to the reference's stable (tile, depth) order) and, in the
//      same pass, gathers the 48-byte records into depth-sorted, per-tile contiguous order for the bulk-TMA staging
//      of the render kernels.
#pragma once
#include "dgr_common.cuh"

namespace dgr {

// binning scratch: [keys u64 x cap][ids u32 x cap][rec_sorted 48 x cap]
struct BinningLayout {
    size_t off_keys, off_ids, off_rec, total;
    __host__ explicit BinningLayout(uint64_t cap) {
        size_t c = cap > 0 ? (size_t)cap : 1;
        size_t o = 0;
        off_keys = o; o = align_up(o + c * 8, 256);
        off_ids = o;  o = align_up(o + c * 4, 256);
        off_rec = o;  o = align_up(o + c * sizeof(Rec), 256);
        total = o;
    }
};

// Column scan of the [nblocks x tiles] count matrix (in place -> exclusive per-(block, tile) offsets) + column totals.
// One CTA = 32 tiles (lanes, coalesced 128-byte rows) x 32 block-groups (warps): every thread first sums its slice of
// the column, the 32 partial sums are scanned through shared memory, then the slice is rewritten as running offsets.
__global__ void __launch_bounds__(1024)
tile_colscan_kernel(int tiles, int nblocks, unsigned *__restrict__ blk_hist, unsigned *__restrict__ tile_count) {
    __shared__ unsigned s_part[32][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const int per = (nblocks + 31) / 32;
    const int b0 = grp * per, b1 = min(nblocks, b0 + per);
    unsigned sum = 0;
    if (t < tiles) {
        const unsigned *p = blk_hist + t;
#pragma unroll 4
        for (int b = b0; b < b1; b++) sum += p[(size_t)b * tiles];
    }
    s_part[grp][lane] = sum;
    __syncthreads();
    if (grp == 0) {                       // one warp: exclusive scan over the 32 groups of each tile (lane = tile)
        unsigned run = 0;
#pragma unroll
        for (int g = 0; g < 32; g++) { const unsigned v = s_part[g][lane]; s_part[g][lane] = run; run += v; }
        if (t < tiles) tile_count[t] = run;
    }
    __syncthreads();
    if (t < tiles) {
        unsigned run = s_part[grp][lane];
        unsigned *p = blk_hist + t;
#pragma unroll 4
        for (int b = b0; b < b1; b++) { const unsigned c = p[(size_t)b * tiles]; p[(size_t)b * tiles] = run; run += c; }
    }
}

__device__ __forceinline__ void cmpex(unsigned long long *a, int i, int l) {
    const unsigned long long x = a[i], y = a[l];
    if (x > y) { a[i] = y; a[l] = x; }
}

// Ascending-only bitonic network for arbitrary n (comparators whose upper index is >= n are skipped: they would
// compare against a virtual +inf that never moves).  All index arithmetic is shift/mask.
__device__ __forceinline__ void bitonic_sort_cta(unsigned long long *a, int n) {
    int lg = 0;
    while ((1 << lg) < n) lg++;
    const int half = (1 << lg) >> 1;
    for (int lk = 1; lk <= lg; lk++) {
        const int k = 1 << lk, hk = k >> 1;
        for (int t = threadIdx.x; t < half; t += blockDim.x) {
            const int blk = t >> (lk - 1), off = t & (hk - 1);
            const int i = (blk << lk) + off, l = (blk << lk) + (k - 1 - off);
            if (l < n) cmpex(a, i, l);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i + j;
                if (l < n) cmpex(a, i, l);
            }
            __syncthreads();
        }
    }
}

// Sort-kernel populations: tiles with n <= kSortSmallCap are sorted by the grid-wide kernel (one CTA per tile);
// larger ones are appended to a compact list that a small persistent grid walks.
constexpr int kSortSmallCap = 4096;
constexpr int kOrderBins = 128;          // log-scale population classes for the heaviest-first issue order

struct TileWork {                        // written by tile_scan_kernel (lives in image scratch)
    unsigned n_big;                      // number of entries of big_list
    unsigned pad[3];
};

__device__ __forceinline__ int order_bin(unsigned count) {
    // descending population class: 0 = heaviest.  class = 4 * floor(log2(count)) + next two mantissa bits
    if (count == 0) return kOrderBins - 1;
    const int lg = 31 - __clz(count);
    const int frac = lg >= 2 ? (int)((count >> (lg - 2)) & 3u) : (int)((count << (2 - lg)) & 3u);
    const int cls = lg * 4 + frac;                      // 0 .. 127
    return max(0, kOrderBins - 2 - cls);
}

// One CTA: exclusive scan of the per-tile instance counts -> ranges (clipped to `cap`), total -> header; plus the
// heaviest-first issue order of the render kernels (counting sort on log-scale population classes — longest
// processing time first keeps the big tiles off the tail) and the list of tiles too big for the per-tile sort CTA.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int tiles, const unsigned *__restrict__ tile_count, unsigned long long cap, uint2 *__restrict__ ranges,
                 GeomHeader *__restrict__ hdr, unsigned *__restrict__ tile_order, TileWork *__restrict__ work,
                 unsigned *__restrict__ big_list) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry, s_total;
    __shared__ unsigned s_bin[kOrderBins], s_nbig;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; s_nbig = 0; }
    if (tid < kOrderBins) s_bin[tid] = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int t = base + tid;
        const unsigned long long v = (t < tiles) ? tile_count[t] : 0u;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
        if (lane == 31) s_warp[warp] = inc;                              // warp totals
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = s_warp[lane];
            unsigned long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long n = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += n; }
            s_warp[lane] = winc - w;                                     // exclusive warp offsets
            if (lane == 31) s_total = winc;
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[warp] + (inc - v);
        if (t < tiles) {
            const unsigned long long s = excl < cap ? excl : cap;
            const unsigned long long e = (excl + v) < cap ? (excl + v) : cap;
            ranges[t] = make_uint2((unsigned)s, (unsigned)e);
            atomicAdd(&s_bin[order_bin((unsigned)(e - s))], 1u);
            if ((unsigned)(e - s) > (unsigned)kSortSmallCap) big_list[atomicAdd(&s_nbig, 1u)] = (unsigned)t;
        }
        __syncthreads();
        if (tid == 0) s_carry += s_total;
        __syncthreads();
    }
    if (tid == 0) { hdr->n_inst = s_carry; work->n_big = s_nbig; }
    // counting sort of the tiles by population class
    if (warp == 0) {
        unsigned run = 0;
        for (int b0 = 0; b0 < kOrderBins; b0 += 32) {
            const unsigned v = s_bin[b0 + lane];
            unsigned inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
            s_bin[b0 + lane] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int t = tid; t < tiles; t += 1024) {
        const uint2 r = ranges[t];
        tile_order[atomicAdd(&s_bin[order_bin(r.y - r.x)], 1u)] = (unsigned)t;
    }
}

// Same block <-> Gaussian mapping as the preprocess kernel.  s_off[t] = ranges[t].start + offset[block][t];
// every instance takes the next slot of its tile with a shared-memory atomic.
__global__ void __launch_bounds__(kPreThreads)
emit_instances_kernel(int P, int gx, int tiles, int gpb_iters, const Rec *__restrict__ rec, const unsigned *__restrict__ touched,
                      const uint2 *__restrict__ ranges, const unsigned *__restrict__ blk_off,
                      unsigned long long *__restrict__ keys) {
    extern __shared__ unsigned s_off[];
    const unsigned *row = blk_off + (size_t)blockIdx.x * tiles;
    for (int t = threadIdx.x; t < tiles; t += kPreThreads) s_off[t] = ranges[t].x + row[t];
    __syncthreads();
    for (int it = 0; it < gpb_iters; it++) {
        const int g = (int)((blockIdx.x * gpb_iters + it) * kPreThreads + threadIdx.x);
        if (g >= P) continue;
        if (touched[g] == 0) continue;
        const float4 q1 = rec[g].q1;
        const unsigned ax = __float_as_uint(q1.w), ay = __float_as_uint(rec[g].q2.w);
        const int tx0 = (int)(ax & 0xffffu) >> 4, tx1 = (int)(ax >> 16) >> 4;
        const int ty0 = (int)(ay & 0xffffu) >> 4, ty1 = (int)(ay >> 16) >> 4;
        const unsigned long long key = ((unsigned long long)__float_as_uint(q1.z) << 32) | (unsigned)g;
        for (int y = ty0; y <= ty1; y++)
            for (int x = tx0; x <= tx1; x++) {
                const int t = y * gx + x;
                const unsigned pos = atomicAdd(&s_off[t], 1u);
                if (pos < __ldg(&ranges[t].y)) keys[pos] = key;           // beyond the (clipped) range: dropped
            }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Per-tile sort by (depth, Gaussian id) + record gather.
//
// Sorting 64-bit keys with a comparison network costs O(n log^2 n) shared-memory compare-exchanges; instead each tile
// quantises its depths to 16 bits over ITS OWN [min, max] depth range (monotone), runs a stable 2-pass LSD radix sort
// (8-bit digits, warp match.any ranking) on 32-bit (q16 << 16 | index) items, and then repairs the few runs of equal
// q16 by comparing the full 64-bit keys.  The result is exactly the (depth, id) order — the reference's stable
// (tile, depth) order — at ~100 instructions per key.  Degenerate inputs (a long run of equal quantised depth, e.g. a
// planar cloud seen head-on) fall back to the bitonic network on the full keys.
// ----------------------------------------------------------------------------------------------------------------
template <int THREADS, int CAP>
struct SortSmem {
    static constexpr int kWarps = THREADS / 32;
    static constexpr size_t off_keys = 0;                                   // u64 [CAP]
    static constexpr size_t off_i0 = off_keys + (size_t)CAP * 8;            // u32 [CAP]
    static constexpr size_t off_i1 = off_i0 + (size_t)CAP * 4;              // u32 [CAP]
    static constexpr size_t off_cnt = off_i1 + (size_t)CAP * 4;             // u16 [kWarps][256]
    static constexpr size_t off_base = off_cnt + (size_t)kWarps * 256 * 2;  // u32 [256]
    static constexpr size_t off_misc = off_base + 256 * 4;                  // u32 [8 + kWarps * 2]
    static constexpr size_t bytes = off_misc + (8 + kWarps * 2) * 4;
};

template <int THREADS>
__device__ __forceinline__ void radix_pass(const unsigned *__restrict__ src, unsigned *__restrict__ dst, int n, int shift,
                                           unsigned short *cnt /*[W][256]*/, unsigned *base /*[256]*/) {
    constexpr int W = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = ((n + W - 1) / W + 31) & ~31;                          // items per warp, multiple of 32
    const int w0 = warp * seg, w1 = min(n, w0 + seg);
    for (int i = tid; i < W * 256; i += THREADS) cnt[i] = 0;
    __syncthreads();
    unsigned short *mycnt = cnt + warp * 256;
    for (int e = w0 + lane; e - lane < w1; e += 32) {                       // warp-uniform trip count
        const bool valid = e < w1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned d = (src[e] >> shift) & 0xffu;
            const unsigned peers = __match_any_sync(act, d);
            if ((unsigned)lane == (unsigned)(__ffs(peers) - 1)) mycnt[d] = (unsigned short)(mycnt[d] + __popc(peers));
        }
        __syncwarp();
    }
    __syncthreads();
    // digit totals and per-warp exclusive prefixes: thread d (< 256) owns digit d
    unsigned total = 0;
    if (tid < 256) {
#pragma unroll 4
        for (int w = 0; w < W; w++) { const unsigned c = cnt[w * 256 + tid]; cnt[w * 256 + tid] = (unsigned short)total; total += c; }
    }
    // exclusive scan of the 256 totals (threads 0..255 = 8 warps)
    unsigned inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    __shared__ unsigned s_wsum[8];
    if (tid < 256 && lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (tid < 256) {
        unsigned off = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w < warp) off += s_wsum[w];
        base[tid] = off + inc - total;
    }
    __syncthreads();
    for (int e = w0 + lane; e - lane < w1; e += 32) {
        const bool valid = e < w1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        unsigned d = 0, peers = 0, item = 0, pos = 0;
        if (valid) {
            item = src[e];
            d = (item >> shift) & 0xffu;
            peers = __match_any_sync(act, d);
            pos = base[d] + mycnt[d] + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
        if (valid) {
            dst[pos] = item;
            if ((unsigned)lane == (unsigned)(__ffs(peers) - 1)) mycnt[d] = (unsigned short)(mycnt[d] + __popc(peers));
        }
        __syncwarp();
    }
    __syncthreads();
}

template <int THREADS, int CAP>
__device__ __forceinline__ void sort_gather_tile(const uint2 r, unsigned long long *__restrict__ keys, const Rec *__restrict__ rec,
                                                 unsigned *__restrict__ ids_sorted, Rec *__restrict__ rec_sorted, unsigned char *smem) {
    using SM = SortSmem<THREADS, CAP>;
    const int n = (int)(r.y - r.x);
    const int tid = threadIdx.x;
    unsigned long long *gk = keys + r.x;
    if (n > CAP) {                                     // beyond the shared-memory capacity: in-place network in global memory
        __syncthreads();
        bitonic_sort_cta(gk, n);
        for (int i = tid; i < n; i += THREADS) {
            const unsigned gid = (unsigned)(gk[i] & 0xffffffffull);
            ids_sorted[r.x + i] = gid;
            const Rec *src = rec + gid;
            Rec v; v.q0 = __ldg(&src->q0); v.q1 = __ldg(&src->q1); v.q2 = __ldg(&src->q2);
            rec_sorted[r.x + i] = v;
        }
        __syncthreads();
        return;
    }
    unsigned long long *A = reinterpret_cast<unsigned long long *>(smem + SM::off_keys);
    unsigned *I0 = reinterpret_cast<unsigned *>(smem + SM::off_i0), *I1 = reinterpret_cast<unsigned *>(smem + SM::off_i1);
    unsigned short *cnt = reinterpret_cast<unsigned short *>(smem + SM::off_cnt);
    unsigned *base = reinterpret_cast<unsigned *>(smem + SM::off_base);
    unsigned *misc = reinterpret_cast<unsigned *>(smem + SM::off_misc);     // [0] min bits, [1] max bits, [2] long-run flag
    if (tid == 0) { misc[0] = 0xffffffffu; misc[1] = 0u; misc[2] = 0u; }
    __syncthreads();
    unsigned lo = 0xffffffffu, hi = 0u;
    for (int i = tid; i < n; i += THREADS) {
        const unsigned long long k = gk[i];
        A[i] = k;
        const unsigned d = (unsigned)(k >> 32);
        lo = min(lo, d); hi = max(hi, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((tid & 31) == 0) { atomicMin(&misc[0], lo); atomicMax(&misc[1], hi); }
    __syncthreads();
    if (n > 1) {
        const float dmin = __uint_as_float(misc[0]), dmax = __uint_as_float(misc[1]);
        const float scale = (dmax > dmin) ? 65535.f / (dmax - dmin) : 0.f;
        for (int i = tid; i < n; i += THREADS) {
            const float d = __uint_as_float((unsigned)(A[i] >> 32));
            const unsigned q = min(65535u, (unsigned)((d - dmin) * scale));          // monotone in d
            I0[i] = (q << 16) | (unsigned)i;
        }
        __syncthreads();
        radix_pass<THREADS>(I0, I1, n, 16, cnt, base);
        radix_pass<THREADS>(I1, I0, n, 24, cnt, base);
        // repair runs of equal q16 with the full (depth, id) keys; each run is owned by the thread of its first element
        for (int i = tid; i < n; i += THREADS) {
            const unsigned q = I0[i] >> 16;
            if ((i == 0 || (I0[i - 1] >> 16) != q) && (i + 1 < n) && (I0[i + 1] >> 16) == q) {
                int e = i + 2;
                while (e < n && (I0[e] >> 16) == q && e - i <= 16) e++;
                if (e - i > 16) { misc[2] = 1u; }
                else {
                    for (int a = i + 1; a < e; a++) {                                  // insertion sort of the run
                        const unsigned it = I0[a];
                        const unsigned long long ka = A[it & 0xffffu];
                        int b = a - 1;
                        while (b >= i && A[I0[b] & 0xffffu] > ka) { I0[b + 1] = I0[b]; b--; }
                        I0[b + 1] = it;
                    }
                }
            }
        }
        __syncthreads();
        if (misc[2]) {                                  // degenerate depth distribution: comparison network on the full keys
            bitonic_sort_cta(A, n);
            for (int i = tid; i < n; i += THREADS) I0[i] = (unsigned)i;
            __syncthreads();
        }
    } else {
        if (tid == 0) I0[0] = 0u;
        __syncthreads();
    }
    for (int i = tid; i < n; i += THREADS) {
        const unsigned gid = (unsigned)(A[I0[i] & 0xffffu] & 0xffffffffull);
        ids_sorted[r.x + i] = gid;
        const Rec *src = rec + gid;
        Rec v; v.q0 = __ldg(&src->q0); v.q1 = __ldg(&src->q1); v.q2 = __ldg(&src->q2);
        rec_sorted[r.x + i] = v;
    }
    __syncthreads();
}

constexpr int kSortSmallThreads = 256;
constexpr int kSortBigThreads = 1024, kSortBigCap = 11264;     // 176 KB keys+items + 16 KB counters

// grid = tiles: one CTA per tile with 0 < n <= kSortSmallCap
__global__ void __launch_bounds__(kSortSmallThreads)
tile_sort_gather_kernel(const unsigned *__restrict__ tile_order, const uint2 *__restrict__ ranges, unsigned long long *__restrict__ keys,
                        const Rec *__restrict__ rec, unsigned *__restrict__ ids_sorted, Rec *__restrict__ rec_sorted) {
    extern __shared__ __align__(16) unsigned char s_sort[];
    const uint2 r = ranges[tile_order[blockIdx.x]];
    const int n = (int)(r.y - r.x);
    if (n <= 0 || n > kSortSmallCap) return;
    sort_gather_tile<kSortSmallThreads, kSortSmallCap>(r, keys, rec, ids_sorted, rec_sorted, s_sort);
}

// small persistent grid walking the list of big tiles
__global__ void __launch_bounds__(kSortBigThreads)
tile_sort_gather_big_kernel(const TileWork *__restrict__ work, const unsigned *__restrict__ big_list, const uint2 *__restrict__ ranges,
                            unsigned long long *__restrict__ keys, const Rec *__restrict__ rec, unsigned *__restrict__ ids_sorted,
                            Rec *__restrict__ rec_sorted) {
    extern __shared__ __align__(16) unsigned char s_sort[];
    const unsigned nb = work->n_big;
    for (unsigned i = blockIdx.x; i < nb; i += gridDim.x)
        sort_gather_tile<kSortBigThreads, kSortBigCap>(ranges[big_list[i]], keys, rec, ids_sorted, rec_sorted, s_sort);
}

}  // namespace dgr
