// dgr_render.cuh — per-tile front-to-back compositing (A4) and its reverse-traversal backward (A5).
//
// B200 design (not the reference's one-thread-per-pixel cooperative fetch with a block barrier per batch):
//   * a tile's depth-sorted 48-byte records are one contiguous block.  A dedicated PRODUCER warp streams it into a
//     4-stage shared-memory ring with 1-D bulk TMA (cp.async.bulk -> UBLKCP); "full" mbarriers carry the transaction
//     bytes, "empty" mbarriers (one arrival per consumer warp) hand a stage back.  There is no block-wide barrier in the
//     main loop: every consumer warp runs through the tile's list at its own pace (up to 3 chunks ahead of the slowest);
//   * each consumer warp owns a sub-tile of 32*PPL pixels (PPL = pixels per lane: 8x4, 8x8 or 16x8).  For every batch of
//     32 staged records each LANE tests ONE record's opacity-aware pixel AABB against the warp's sub-tile, the ballot
//     gives the records that can touch the sub-tile at all, and only those are evaluated (records are broadcast-read
//     from shared memory).  This skips most of the (pixel, Gaussian) pairs the reference evaluates and then discards
//     at alpha < 1/255;
//   * tiles are issued heaviest-first (tile_order from the scan kernel), so the long tiles do not form the tail;
//   * backward: per (warp, record) the 10 partial sums are added over the lane's PPL pixels, reduced over the warp with
//     a 13-shuffle recursive-halving butterfly (warp-shuffle reduction) and land on 10 lanes which issue ONE
//     red.global.add each — instead of the reference's ~10 atomics per (pixel, Gaussian) pair.
// Results follow the reference's rules exactly: power > 0 skip, alpha = min(0.99, o G), alpha < 1/255 skip,
// stop at T (1 - alpha) < 1e-4, colour + T*bg, un-normalised depth, alpha = sum alpha T.
#pragma once
#include "dgr_common.cuh"

namespace dgr {

constexpr int kChunk = 128;   // records per shared-memory stage (6 KB)
constexpr int kStages = 4;    // bulk-TMA ring depth (24 KB of shared memory per CTA)

// power * log2(e) for pixel offset (dx, dy); identical instruction sequence in forward and backward so both make
// the same skip decisions.   q0.z = -0.5 A log2e, q0.w = -B log2e, q1.x = -0.5 C log2e
__device__ __forceinline__ float eval_power2(const float4 &q0, const float4 &q1, float dx, float dy) {
    const float t = __fmaf_rn(q0.z, dx, __fmul_rn(q0.w, dy));
    return __fmaf_rn(dx, t, __fmul_rn(__fmul_rn(q1.x, dy), dy));
}

__device__ __forceinline__ bool aabb_hit(unsigned ax, unsigned ay, int wx0, int wx1, int wy0, int wy1) {
    const int gx0 = (int)(ax & 0xffffu), gx1 = (int)(ax >> 16), gy0 = (int)(ay & 0xffffu), gy1 = (int)(ay >> 16);
    return (gx0 <= wx1) & (gx1 >= wx0) & (gy0 <= wy1) & (gy1 >= wy0);
}

// Lane-level cull of one staged record against a warp's sub-tile [wx0,wx1] x [wy0,wy1]: the opacity-aware pixel AABB
// (which carries the reference's 3-sigma tile-rect clip) AND the exact "can the alpha >= 1/255 ellipse reach it" test.
__device__ __forceinline__ bool record_hits_subtile(const Rec &r, int wx0, int wx1, int wy0, int wy1) {
    const float4 q1 = r.q1;
    const unsigned ax = __float_as_uint(q1.w), ay = __float_as_uint(r.q2.w);
    const int gx0 = (int)(ax & 0xffffu), gx1 = (int)(ax >> 16), gy0 = (int)(ay & 0xffffu), gy1 = (int)(ay >> 16);
    const int x0 = max(gx0, wx0), x1 = min(gx1, wx1), y0 = max(gy0, wy0), y1 = min(gy1, wy1);
    if (x0 > x1 || y0 > y1) return false;
    const float4 q0 = r.q0;
    return ellipse_hits_rect(q0.x, q0.y, q0.z, q0.w, q1.x, alpha_threshold_power2(q1.y), (float)x0, (float)x1, (float)y0, (float)y1);
}

// Sub-tile geometry of one consumer warp for PPL pixels per lane.
template <int PPL>
struct SubTile {
    static constexpr int kWarps = 8 / PPL;                 // warps per tile (all of them consumers; warp 0 also feeds the ring)
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kW = (PPL == 4) ? 16 : 8;         // region width
    static constexpr int kH = (PPL == 1) ? 4 : 8;          // region height
    __device__ static __forceinline__ int x0(int tx, int warp) { return tx * kTile + ((PPL == 4) ? 0 : (warp & 1) * 8); }
    __device__ static __forceinline__ int y0(int ty, int warp) { return ty * kTile + ((PPL == 4) ? warp * 8 : (warp >> 1) * kH); }
    __device__ static __forceinline__ int px(int wx0, int lane, int p) { return wx0 + (lane & 7) + ((PPL == 4 && (p & 1)) ? 8 : 0); }
    __device__ static __forceinline__ int py(int wy0, int lane, int p) { return wy0 + (lane >> 3) + ((PPL == 4) ? (p >> 1) * 4 : p * 4); }
};

struct RingSmem {
    Rec rec[kStages][kChunk];
    uint64_t full[kStages], empty[kStages];
    unsigned done_warps;      // forward: consumer warps whose pixels are all finished
    unsigned maxlast;         // backward: longest per-pixel list of the tile
};

template <int PPL>
__global__ void __launch_bounds__(SubTile<PPL>::kThreads)
render_fwd_kernel(int H, int W, int gx, const unsigned *__restrict__ tile_order, const uint2 *__restrict__ ranges,
                  const Rec *__restrict__ rec_sorted, const float *__restrict__ bg, float *__restrict__ out_color,
                  float *__restrict__ out_depth, float *__restrict__ out_alpha, unsigned *__restrict__ n_contrib,
                  float *__restrict__ final_T) {
    using ST = SubTile<PPL>;
    constexpr unsigned NW = ST::kWarps;
    __shared__ __align__(128) RingSmem sm;
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nchunks = (n + kChunk - 1) / kChunk;
    const Rec *src = rec_sorted + range.x;
    volatile unsigned *vdone = &sm.done_warps;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; i++) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], NW); }
        sm.done_warps = 0;
        mbar_fence_init();
    }
    __syncthreads();

    // Warp 0's lane 0 feeds the ring: at the top of chunk c it makes sure chunk c itself is in flight (blocking on the
    // stage's "empty" barrier if it must) and then prefetches up to kStages - 1 chunks ahead without blocking.
    int next_issue = 0;
    auto issue = [&](int c) {
        const int st = c % kStages;
        const uint32_t bytes = (uint32_t)min(kChunk, n - c * kChunk) * (uint32_t)sizeof(Rec);
        mbar_expect_tx(&sm.full[st], bytes);
        tma_bulk_g2s(&sm.rec[st][0], src + (size_t)c * kChunk, bytes, &sm.full[st]);
    };
    auto produce = [&](int c) {          // returns false when every warp of the tile is finished
        while (next_issue < nchunks && next_issue <= c) {
            if (next_issue >= kStages) {
                const uint32_t par = (uint32_t)(((next_issue / kStages) - 1) & 1);
                while (!mbar_try_wait(&sm.empty[next_issue % kStages], par)) { if (*vdone == NW) return false; }
            }
            issue(next_issue++);
        }
        while (next_issue < nchunks && next_issue < c + kStages) {
            if (next_issue >= kStages && !mbar_try_wait(&sm.empty[next_issue % kStages], (uint32_t)(((next_issue / kStages) - 1) & 1))) break;
            issue(next_issue++);
        }
        return true;
    };

    const int wx0 = ST::x0(tx, warp), wy0 = ST::y0(ty, warp);
    const int wx1 = wx0 + ST::kW - 1, wy1 = wy0 + ST::kH - 1;
    float fx[PPL], fy[PPL], T[PPL], C0[PPL], C1[PPL], C2[PPL], D[PPL];
    unsigned last[PPL];
    bool done[PPL], inside[PPL];
    bool all_done = true;
#pragma unroll
    for (int p = 0; p < PPL; p++) {
        const int x = ST::px(wx0, lane, p), y = ST::py(wy0, lane, p);
        inside[p] = (x < W) && (y < H);
        fx[p] = (float)x; fy[p] = (float)y;
        T[p] = 1.f; C0[p] = 0.f; C1[p] = 0.f; C2[p] = 0.f; D[p] = 0.f; last[p] = 0; done[p] = !inside[p];
        all_done = all_done && done[p];
    }
    bool warp_done = __all_sync(0xffffffffu, all_done);
    if (warp_done && lane == 0) atomicAdd(&sm.done_warps, 1u);
    for (int c = 0; c < nchunks; c++) {
        const int s = c % kStages;
        if (warp == 0) {
            bool go = true;
            if (lane == 0) go = produce(c);
            if (!__shfl_sync(0xffffffffu, go ? 1 : 0, 0)) break;
        }
        if (!warp_done) {
            mbar_wait(&sm.full[s], (uint32_t)((c / kStages) & 1));
            const int cnt = min(kChunk, n - c * kChunk);
            for (int b = 0; b < cnt; b += 32) {
                const int i = b + lane;
                bool hit = false;
                if (i < cnt) hit = record_hits_subtile(sm.rec[s][i], wx0, wx1, wy0, wy1);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const Rec *r = &sm.rec[s][b + j];
                    const float4 q0 = r->q0, q1 = r->q1;
                    float4 q2;
                    bool have_q2 = false;
#pragma unroll
                    for (int p = 0; p < PPL; p++) {
                        const float dx = q0.x - fx[p], dy = q0.y - fy[p];
                        const float p2 = eval_power2(q0, q1, dx, dy);
                        const float ag = __fmul_rn(q1.y, ex2_approx(p2));
                        const float a = fminf(DGR_ALPHA_MAX, ag);
                        bool ok = (!done[p]) & (p2 <= 0.f) & (a >= DGR_ALPHA_MIN);
                        const float test_T = __fmul_rn(T[p], 1.f - a);
                        if (ok && test_T < DGR_T_STOP) { done[p] = true; ok = false; }
                        if (ok) {
                            if (!have_q2) { q2 = r->q2; have_q2 = true; }
                            const float w = __fmul_rn(a, T[p]);
                            C0[p] = __fmaf_rn(q2.x, w, C0[p]); C1[p] = __fmaf_rn(q2.y, w, C1[p]); C2[p] = __fmaf_rn(q2.z, w, C2[p]);
                            D[p] = __fmaf_rn(q1.z, w, D[p]);
                            T[p] = test_T;
                            last[p] = (unsigned)(c * kChunk + b + j + 1);
                        }
                    }
                }
                all_done = true;
#pragma unroll
                for (int p = 0; p < PPL; p++) all_done = all_done && done[p];
                if (__all_sync(0xffffffffu, all_done)) { warp_done = true; break; }
            }
            __syncwarp();
            if (lane == 0) {
                if (warp_done) atomicAdd(&sm.done_warps, 1u);
                mbar_arrive(&sm.empty[s]);
            }
        } else {
            // finished warp: keep handing stages back (in phase order) until every warp of the tile is finished
            if (*vdone == NW) break;
            bool stop = false;
            if (c >= kStages) {
                const uint32_t par = (uint32_t)(((c / kStages) - 1) & 1);
                while (!mbar_try_wait(&sm.empty[s], par)) { if (*vdone == NW) { stop = true; break; } }
            }
            if (stop) break;
            if (lane == 0) mbar_arrive(&sm.empty[s]);
            __syncwarp();
        }
    }
    // never leave the CTA with a bulk copy in flight into its shared memory
    if (warp == 0 && lane == 0)
        for (int cc = max(0, next_issue - kStages); cc < next_issue; cc++) mbar_wait(&sm.full[cc % kStages], (uint32_t)((cc / kStages) & 1));

    const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int p = 0; p < PPL; p++) {
        if (inside[p]) {
            const size_t pix = (size_t)ST::py(wy0, lane, p) * W + ST::px(wx0, lane, p);
            out_color[pix] = C0[p] + T[p] * b0;
            out_color[HW + pix] = C1[p] + T[p] * b1;
            out_color[2 * HW + pix] = C2[p] + T[p] * b2;
            out_depth[pix] = D[p];
            out_alpha[pix] = 1.f - T[p];
            n_contrib[pix] = last[p];
            final_T[pix] = T[p];
        }
    }
}

// 12 values per lane -> one value per lane; lane L ends with component
//   comp(L) = 6*b4 + 3*b3 + {b2b1: 00->0, 01->1, 10->2, 11->none}   (b0 duplicates)
__device__ __forceinline__ float reduce12(const float (&v)[12], int lane) {
    float a[6], b[3];
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const float keep = hi ? v[i + 6] : v[i], send = hi ? v[i] : v[i + 6];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float keep = hi ? a[i + 3] : a[i], send = hi ? a[i] : a[i + 3];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = (lane & 4) != 0;
    const float c0 = (hi ? b[2] : b[0]) + __shfl_xor_sync(0xffffffffu, hi ? b[0] : b[2], 4);
    const float c1 = (hi ? 0.f : b[1]) + __shfl_xor_sync(0xffffffffu, hi ? b[1] : 0.f, 4);
    hi = (lane & 2) != 0;
    float d = (hi ? c1 : c0) + __shfl_xor_sync(0xffffffffu, hi ? c0 : c1, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// two independent 12-value reductions with their shuffles interleaved (twice the shuffles in flight per dependent round)
__device__ __forceinline__ void reduce12x2(const float (&v)[12], const float (&w)[12], int lane, float &rv, float &rw) {
    float a[6], b[3], e[6], f[3];
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const float keep = hi ? v[i + 6] : v[i], send = hi ? v[i] : v[i + 6];
        const float keep2 = hi ? w[i + 6] : w[i], send2 = hi ? w[i] : w[i + 6];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        e[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float keep = hi ? a[i + 3] : a[i], send = hi ? a[i] : a[i + 3];
        const float keep2 = hi ? e[i + 3] : e[i], send2 = hi ? e[i] : e[i + 3];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        f[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 8);
    }
    hi = (lane & 4) != 0;
    const float c0 = (hi ? b[2] : b[0]) + __shfl_xor_sync(0xffffffffu, hi ? b[0] : b[2], 4);
    const float g0 = (hi ? f[2] : f[0]) + __shfl_xor_sync(0xffffffffu, hi ? f[0] : f[2], 4);
    const float c1 = (hi ? 0.f : b[1]) + __shfl_xor_sync(0xffffffffu, hi ? b[1] : 0.f, 4);
    const float g1 = (hi ? 0.f : f[1]) + __shfl_xor_sync(0xffffffffu, hi ? f[1] : 0.f, 4);
    hi = (lane & 2) != 0;
    float d = (hi ? c1 : c0) + __shfl_xor_sync(0xffffffffu, hi ? c0 : c1, 2);
    float h = (hi ? g1 : g0) + __shfl_xor_sync(0xffffffffu, hi ? g0 : g1, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    h += __shfl_xor_sync(0xffffffffu, h, 1);
    rv = d; rw = h;
}

template <int PPL, bool U2>
__global__ void __launch_bounds__(SubTile<PPL>::kThreads)
render_bwd_kernel(int H, int W, int gx, const unsigned *__restrict__ tile_order, const uint2 *__restrict__ ranges,
                  const Rec *__restrict__ rec_sorted, const unsigned *__restrict__ ids_sorted, const float *__restrict__ bg,
                  const float *__restrict__ final_T, const unsigned *__restrict__ n_contrib,
                  const float *__restrict__ gC, const float *__restrict__ gD, const float *__restrict__ gA,
                  float *__restrict__ grad_rec) {
    using ST = SubTile<PPL>;
    constexpr unsigned NW = ST::kWarps;
    __shared__ __align__(128) RingSmem sm;
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint2 range = ranges[tile];
    if (range.y == range.x) return;
    const int wx0 = ST::x0(tx, warp), wy0 = ST::y0(ty, warp);
    const int wx1 = wx0 + ST::kW - 1, wy1 = wy0 + ST::kH - 1;

    float fx[PPL], fy[PPL], gc0[PPL], gc1[PPL], gc2[PPL], gd[PPL], ga[PPL], T[PPL], R[PPL];
    unsigned last[PPL];
    unsigned lmax = 0;
    const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int p = 0; p < PPL; p++) {
        const int x = ST::px(wx0, lane, p), y = ST::py(wy0, lane, p);
        fx[p] = (float)x; fy[p] = (float)y;
        gc0[p] = 0.f; gc1[p] = 0.f; gc2[p] = 0.f; gd[p] = 0.f; ga[p] = 0.f; T[p] = 1.f; last[p] = 0;
        if ((x < W) && (y < H)) {
            const size_t pix = (size_t)y * W + x;
            last[p] = n_contrib[pix];
            T[p] = __ldg(final_T + pix);
            if (gC) { gc0[p] = __ldg(gC + pix); gc1[p] = __ldg(gC + HW + pix); gc2[p] = __ldg(gC + 2 * HW + pix); }
            if (gD) gd[p] = __ldg(gD + pix);
            if (gA) ga[p] = __ldg(gA + pix);
        }
        // R = T_final * (bg . gC) + sum over Gaussians behind the current one of w * s
        R[p] = T[p] * (b0 * gc0[p] + b1 * gc1[p] + b2 * gc2[p]);
        lmax = max(lmax, last[p]);
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; i++) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], NW); }
        sm.maxlast = 0;
        mbar_fence_init();
    }
    __syncthreads();
    unsigned wmax = lmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0 && wmax) atomicMax(&sm.maxlast, wmax);
    __syncthreads();
    const int n = (int)sm.maxlast;                      // records [0, n) of this tile's list matter
    const int nchunks = (n + kChunk - 1) / kChunk;
    const Rec *src = rec_sorted + range.x;
    const unsigned *ids = ids_sorted + range.x;
    const unsigned wlast = wmax;                        // warp-level bound

    // step k (0 .. nchunks-1) handles chunk c = nchunks-1-k (back to front) in stage k % kStages.  Warp 0's lane 0 feeds
    // the ring: step k itself (blocking if its stage is still in use), then up to kStages - 1 steps ahead (non-blocking).
    int next_issue = 0;
    auto issue = [&](int k) {
        const int c = nchunks - 1 - k, st = k % kStages;
        const uint32_t bytes = (uint32_t)min(kChunk, n - c * kChunk) * (uint32_t)sizeof(Rec);
        mbar_expect_tx(&sm.full[st], bytes);
        tma_bulk_g2s(&sm.rec[st][0], src + (size_t)c * kChunk, bytes, &sm.full[st]);
    };
    auto produce = [&](int k) {
        while (next_issue < nchunks && next_issue <= k) {
            if (next_issue >= kStages) mbar_wait(&sm.empty[next_issue % kStages], (uint32_t)(((next_issue / kStages) - 1) & 1));
            issue(next_issue++);
        }
        while (next_issue < nchunks && next_issue < k + kStages) {
            if (next_issue >= kStages && !mbar_try_wait(&sm.empty[next_issue % kStages], (uint32_t)(((next_issue / kStages) - 1) & 1))) break;
            issue(next_issue++);
        }
    };
    for (int k = 0; k < nchunks; k++) {
        const int c = nchunks - 1 - k, s = k % kStages;
        if (warp == 0) { if (lane == 0) produce(k); __syncwarp(); }
        // Every warp observes EVERY phase of the full barrier, also for chunks it does not need: a warp that skipped the
        // wait could come back to this stage while the barrier is still one phase behind and the parity test would alias.
        mbar_wait(&sm.full[s], (uint32_t)((k / kStages) & 1));
        if ((unsigned)(c * kChunk) >= wlast) {
            // nothing of this chunk reaches this warp's pixels: hand the stage back without touching the data
            if (lane == 0) mbar_arrive(&sm.empty[s]);
            __syncwarp();
            continue;
        }
        const int cnt = min(kChunk, n - c * kChunk);
        for (int b = ((cnt - 1) >> 5) << 5; b >= 0; b -= 32) {
            if ((unsigned)(c * kChunk + b) >= wlast) continue;
            const int i = b + lane;
            bool hit = false;
            unsigned my_id = 0;
            if (i < cnt) {
                hit = record_hits_subtile(sm.rec[s][i], wx0, wx1, wy0, wy1);
                my_id = __ldg(ids + (size_t)c * kChunk + i);
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            // evaluates record b+j for this warp's pixels, advancing T and R; v = this lane's 10 moment contributions;
            // returns (warp-uniform) whether any lane contributed
            auto eval = [&](int j, float (&v)[12]) -> bool {
                const Rec *r = &sm.rec[s][b + j];
                const float4 q0 = r->q0, q1 = r->q1;
                const unsigned gidx = (unsigned)(c * kChunk + b + j);
                float dxv[PPL], dyv[PPL], agv[PPL], av[PPL];
                bool okv[PPL];
                bool any_ok = false;
#pragma unroll
                for (int p = 0; p < PPL; p++) {
                    dxv[p] = q0.x - fx[p]; dyv[p] = q0.y - fy[p];
                    const float p2 = eval_power2(q0, q1, dxv[p], dyv[p]);
                    agv[p] = __fmul_rn(q1.y, ex2_approx(p2));
                    av[p] = fminf(DGR_ALPHA_MAX, agv[p]);
                    okv[p] = (gidx < last[p]) & (p2 <= 0.f) & (av[p] >= DGR_ALPHA_MIN);
                    any_ok = any_ok || okv[p];
                }
                if (!__any_sync(0xffffffffu, any_ok)) return false;
#pragma unroll
                for (int q = 0; q < 12; q++) v[q] = 0.f;
                if (any_ok) {
                    const float4 q2 = r->q2;
#pragma unroll
                    for (int p = 0; p < PPL; p++) {
                        if (okv[p]) {
                            const float ir = rcp_approx(1.f - av[p]);
                            T[p] = T[p] * ir;
                            const float sdot = __fmaf_rn(q2.x, gc0[p], __fmaf_rn(q2.y, gc1[p], __fmaf_rn(q2.z, gc2[p], __fmaf_rn(q1.z, gd[p], ga[p]))));
                            const float dL_da = T[p] * sdot - R[p] * ir;
                            const float w = av[p] * T[p];
                            R[p] = __fmaf_rn(w, sdot, R[p]);
                            const float u = agv[p] * dL_da;
                            const float udx = u * dxv[p], udy = u * dyv[p];
                            v[0] += u; v[1] += udx; v[2] += udy;
                            v[3] = __fmaf_rn(udx, dxv[p], v[3]); v[4] = __fmaf_rn(udx, dyv[p], v[4]); v[5] = __fmaf_rn(udy, dyv[p], v[5]);
                            v[6] = __fmaf_rn(w, gc0[p], v[6]); v[7] = __fmaf_rn(w, gc1[p], v[7]); v[8] = __fmaf_rn(w, gc2[p], v[8]);
                            v[9] = __fmaf_rn(w, gd[p], v[9]);
                        }
                    }
                }
                return true;
            };
            const int comp = ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0) + ((lane >> 1) & 3);
            const bool writer = ((lane & 1) == 0) && ((lane & 6) != 6) && comp < 10;
            while (mask) {
                const int j = 31 - __clz(mask);
                mask &= ~(1u << j);
                float v[12];
                if (!eval(j, v)) continue;
                if constexpr (U2) {
                    // pair this record with the next contributing one so that two butterflies overlap
                    float v2[12];
                    int j2 = -1;
                    while (mask) {
                        const int jj = 31 - __clz(mask);
                        mask &= ~(1u << jj);
                        if (eval(jj, v2)) { j2 = jj; break; }
                    }
                    if (j2 >= 0) {
                        float red, red2;
                        reduce12x2(v, v2, lane, red, red2);
                        const unsigned gid = __shfl_sync(0xffffffffu, my_id, j), gid2 = __shfl_sync(0xffffffffu, my_id, j2);
                        if (writer) {
                            red_add_f32(grad_rec + (size_t)gid * kGradRecFloats + comp, red);
                            red_add_f32(grad_rec + (size_t)gid2 * kGradRecFloats + comp, red2);
                        }
                        continue;
                    }
                }
                const float red = reduce12(v, lane);
                const unsigned gid = __shfl_sync(0xffffffffu, my_id, j);
                if (writer) red_add_f32(grad_rec + (size_t)gid * kGradRecFloats + comp, red);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
}

}  // namespace dgr
