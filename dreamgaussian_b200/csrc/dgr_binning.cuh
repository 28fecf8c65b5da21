// dgr_binning.cuh — tile binning (A3 of SURVEY.md §8a): one (tile, depth) key + Gaussian id per touched tile,
// stable sort by (tile, depth), per-tile [start, end) ranges, and the gather of the 48-byte records into
// depth-sorted, per-tile contiguous order so the render kernels can stage a tile's block with one bulk-TMA copy.
#pragma once
#include "dgr_common.cuh"

namespace dgr {

// binning scratch: [keys u64 x cap][keys_alt u64 x cap][vals u32 x cap][vals_alt u32 x cap][rec_sorted 48 x cap][sort temp]
struct BinningLayout {
    size_t off_keys, off_keys_alt, off_vals, off_vals_alt, off_rec, off_temp, total;
    __host__ BinningLayout(uint64_t cap, size_t temp_bytes) {
        size_t c = cap > 0 ? (size_t)cap : 1;
        size_t o = 0;
        off_keys = o;     o = align_up(o + c * 8, 256);
        off_keys_alt = o; o = align_up(o + c * 8, 256);
        off_vals = o;     o = align_up(o + c * 4, 256);
        off_vals_alt = o; o = align_up(o + c * 4, 256);
        off_rec = o;      o = align_up(o + c * sizeof(Rec), 256);
        off_temp = o;     o = align_up(o + temp_bytes, 256);
        total = o;
    }
};

// One thread per Gaussian: write its (tile << 32 | depth bits) keys in row-major tile order at offsets[g].
// Instances past `cap` are dropped (the host re-runs with a larger buffer when n_inst > cap).
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, int gx, const Rec *__restrict__ rec, const unsigned *__restrict__ offsets,
                      const unsigned *__restrict__ touched, unsigned long long cap,
                      unsigned long long *__restrict__ keys, unsigned *__restrict__ vals) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    if (touched[g] == 0) return;
    const float4 q1 = rec[g].q1;
    const unsigned ax = __float_as_uint(q1.w), ay = __float_as_uint(rec[g].q2.w);
    const int tx0 = (int)(ax & 0xffffu) >> 4, tx1 = (int)(ax >> 16) >> 4;
    const int ty0 = (int)(ay & 0xffffu) >> 4, ty1 = (int)(ay >> 16) >> 4;
    const unsigned long long depth_bits = __float_as_uint(q1.z);
    unsigned long long off = offsets[g];
    for (int y = ty0; y <= ty1; y++)
        for (int x = tx0; x <= tx1; x++) {
            if (off < cap) {
                keys[off] = ((unsigned long long)(unsigned)(y * gx + x) << 32) | depth_bits;
                vals[off] = (unsigned)g;
            }
            off++;
        }
}

// After the sort: tile ranges from key boundaries + gather records into sorted order.
__global__ void __launch_bounds__(256)
ranges_gather_kernel(unsigned long long n, const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals,
                     const Rec *__restrict__ rec, uint2 *__restrict__ ranges, Rec *__restrict__ rec_sorted) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned tile = (unsigned)(keys[i] >> 32);
    if (i == 0) ranges[tile].x = 0;
    else {
        const unsigned prev = (unsigned)(keys[i - 1] >> 32);
        if (prev != tile) { ranges[prev].y = (unsigned)i; ranges[tile].x = (unsigned)i; }
    }
    if (i == n - 1) ranges[tile].y = (unsigned)n;
    const Rec *src = rec + vals[i];
    Rec r; r.q0 = __ldg(&src->q0); r.q1 = __ldg(&src->q1); r.q2 = __ldg(&src->q2);
    rec_sorted[i] = r;
}

}  // namespace dgr
