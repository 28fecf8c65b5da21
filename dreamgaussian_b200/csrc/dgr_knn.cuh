// dgr_knn.cuh — SURVEY.md §8 row f3: distCUDA2, the mean squared distance of every point to its 3 nearest neighbours
// (reference: /root/reference/simple-knn/simple_knn.cu:185-221 `SimpleKNN::knn`, spatial.cu:15-26 `distCUDA2`; caller
// gs_renderer.py:341).  The result is exact 3-NN, as the reference's (its Morton sort + box pruning only accelerate an
// exact search, simple_knn.cu:132-183); what differs is how the neighbours are found:
//
//   reference: min/max by cub::DeviceReduce + two blocking D2H copies, Morton codes, cub radix sort, thrust temporaries,
//              1024-point boxes tested one by one per point — default stream, host round trips;
//   here:      uniform grid ("cell list") built with a counting sort — bounding box by block reduce + ordered-int atomics,
//              grid shape chosen ON THE DEVICE (about two points per cell, degenerate axes get one cell), count, 3-kernel
//              exclusive scan, scatter into cell order — then one thread per point walks Chebyshev shells of cells until
//              the third-best distance cannot be beaten.  Everything on the caller's stream, no host read-back, no
//              library sort.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace dgr {

struct KnnGrid {                  // device-resident, written by knn_grid_kernel
    unsigned lo[3], hi[3];        // bounding box as order-preserving unsigned encodings of the floats (atomicMin / atomicMax)
    float org[3], inv_h[3], h[3];
    int dim[3];
    int cells;
    float h_min;                  // smallest cell edge over the axes that have more than one cell (FLT_MAX if none)
};

__host__ __device__ inline size_t knn_cells_cap(int P) { size_t c = (size_t)(P > 0 ? P : 1); return c < 4096 ? 4096 : c; }

// scratch: [KnnGrid 256][cell_start u32 x (cap+1)][cell_fill u32 x cap][block sums u32 x nb][cell id u32 x P][sorted float4 x P]
struct KnnLayout {
    size_t off_grid, off_start, off_fill, off_sums, off_cid, off_sorted, total;
    size_t cap, nblk;
    __host__ __device__ explicit KnnLayout(int P) {
        cap = knn_cells_cap(P);
        nblk = (cap + 1 + 4095) / 4096;
        size_t Pn = P > 0 ? (size_t)P : 1, o = 0;
        auto al = [](size_t v) { return (v + 255) / 256 * 256; };
        off_grid = o;   o = al(o + sizeof(KnnGrid));
        off_start = o;  o = al(o + (cap + 1) * 4);
        off_fill = o;   o = al(o + cap * 4);
        off_sums = o;   o = al(o + (nblk + 1) * 4);
        off_cid = o;    o = al(o + Pn * 4);
        off_sorted = o; o = al(o + Pn * 16);
        total = o;
    }
};

__device__ __forceinline__ unsigned knn_enc(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float knn_dec(unsigned e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }

__global__ void __launch_bounds__(256) knn_bbox_kernel(int P, const float *__restrict__ pts, KnnGrid *__restrict__ grid) {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) { const float v = __ldg(pts + 3 * (size_t)i + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicMin(&grid->lo[a], knn_enc(lo[a])); atomicMax(&grid->hi[a], knn_enc(hi[a])); }
    }
}

// One thread: grid shape from the bounding box.  Target: ~2 points per cell, near-cubic cells, at most `cap` cells; an
// axis whose extent is (almost) zero gets a single cell and does not take part in the termination bound.
__global__ void knn_grid_kernel(int P, unsigned cap, KnnGrid *__restrict__ g) {
    float ext[3];
    for (int a = 0; a < 3; a++) { g->org[a] = knn_dec(g->lo[a]); ext[a] = knn_dec(g->hi[a]) - g->org[a]; if (!(ext[a] > 0.f) || !isfinite(ext[a])) ext[a] = 0.f; }
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    bool live[3]; int nlive = 0;
    for (int a = 0; a < 3; a++) { live[a] = emax > 0.f && ext[a] > 1e-6f * emax; nlive += live[a] ? 1 : 0; }
    double target = fmax(1.0, 0.5 * (double)P);
    if (target > (double)cap) target = (double)cap;
    double vol = 1.0;
    for (int a = 0; a < 3; a++) if (live[a]) vol *= (double)ext[a];
    double h = nlive ? pow(vol / target, 1.0 / nlive) : 1.0;
    int dim[3];
    for (int it = 0; it < 64; it++) {                  // grow h until the grid fits the scratch
        double cells = 1.0;
        for (int a = 0; a < 3; a++) { dim[a] = live[a] ? (int)fmin(1024.0, fmax(1.0, ceil((double)ext[a] / h))) : 1; cells *= dim[a]; }
        if (cells <= (double)cap) break;
        h *= 1.15;
    }
    float hmin = FLT_MAX;
    int cells = 1;
    for (int a = 0; a < 3; a++) {
        g->dim[a] = dim[a];
        g->h[a] = dim[a] > 1 ? ext[a] / (float)dim[a] : 0.f;
        g->inv_h[a] = dim[a] > 1 ? (float)dim[a] / ext[a] : 0.f;
        if (dim[a] > 1) hmin = fminf(hmin, g->h[a]);
        cells *= dim[a];
    }
    g->cells = cells;
    g->h_min = hmin * 0.999f;                          // rounding of the cell assignment: stay conservative
}

__device__ __forceinline__ void knn_cell_of(const KnnGrid &g, float x, float y, float z, int &cx, int &cy, int &cz) {
    cx = min(g.dim[0] - 1, max(0, (int)((x - g.org[0]) * g.inv_h[0])));
    cy = min(g.dim[1] - 1, max(0, (int)((y - g.org[1]) * g.inv_h[1])));
    cz = min(g.dim[2] - 1, max(0, (int)((z - g.org[2]) * g.inv_h[2])));
}

__global__ void __launch_bounds__(256)
knn_count_kernel(int P, const float *__restrict__ pts, const KnnGrid *__restrict__ gp, unsigned *__restrict__ cell_count, unsigned *__restrict__ cid) {
    __shared__ KnnGrid g;
    if (threadIdx.x == 0) g = *gp;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int cx, cy, cz;
    knn_cell_of(g, __ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2), cx, cy, cz);
    const unsigned c = (unsigned)((cz * g.dim[1] + cy) * g.dim[0] + cx);
    cid[i] = c;
    atomicAdd(&cell_count[c], 1u);
}

// Exclusive scan of n u32 values in place, 4096 per block: (1) block totals, (2) one block scans the totals, (3) local scan + offset.
__global__ void __launch_bounds__(1024) knn_scan_sums_kernel(const unsigned *__restrict__ v, size_t n, unsigned *__restrict__ sums) {
    __shared__ unsigned s_w[32];
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) if (base + j < n) t += v[base + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned w = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}
__global__ void __launch_bounds__(1024) knn_scan_top_kernel(unsigned *__restrict__ sums, size_t nb) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nb; b0 += 1024) {
        const size_t i = b0 + threadIdx.x;
        const unsigned v = i < nb ? sums[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const unsigned w = s_w[threadIdx.x];
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (threadIdx.x >= o) winc += t; }
            s_w[threadIdx.x] = winc - w;
        }
        __syncthreads();
        const unsigned excl = s_carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) knn_scan_apply_kernel(unsigned *__restrict__ v, size_t n, const unsigned *__restrict__ sums) {
    __shared__ unsigned s_w[32];
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    unsigned x[4], t = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { x[j] = base + j < n ? v[base + j] : 0u; t += x[j]; }
    unsigned inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned w = s_w[threadIdx.x];
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, winc, o); if (threadIdx.x >= o) winc += u; }
        s_w[threadIdx.x] = winc - w;
    }
    __syncthreads();
    unsigned run = sums[blockIdx.x] + s_w[threadIdx.x >> 5] + inc - t;
#pragma unroll
    for (int j = 0; j < 4; j++) { if (base + j < n) v[base + j] = run; run += x[j]; }
}

__global__ void __launch_bounds__(256)
knn_scatter_kernel(int P, const float *__restrict__ pts, const unsigned *__restrict__ cid, const unsigned *__restrict__ cell_start,
                   unsigned *__restrict__ cell_fill, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const unsigned c = cid[i];
    const unsigned pos = cell_start[c] + atomicAdd(&cell_fill[c], 1u);
    sorted[pos] = make_float4(__ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2), __uint_as_float((unsigned)i));
}

// simple_knn.cu:118-130 `updateKBest<3>`: insertion into the ascending triple
__device__ __forceinline__ void knn_update3(float d, float (&best)[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
    }
}

// One thread per point, in CELL order (neighbouring threads search neighbouring cells -> shared cache lines).
__global__ void __launch_bounds__(128)
knn_query_kernel(int P, const KnnGrid *__restrict__ gp, const unsigned *__restrict__ cell_start, const float4 *__restrict__ sorted,
                 float *__restrict__ mean_dists) {
    __shared__ KnnGrid g;
    if (threadIdx.x == 0) g = *gp;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 me = sorted[i];
    int cx, cy, cz;
    knn_cell_of(g, me.x, me.y, me.z, cx, cy, cz);
    float best[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    const int rmax = max(max(max(cx, g.dim[0] - 1 - cx), max(cy, g.dim[1] - 1 - cy)), max(cz, g.dim[2] - 1 - cz));
    auto scan_run = [&](int row, int x0, int x1) {          // cells [x0, x1] of one x-row are contiguous in `sorted`
        const unsigned a = __ldg(cell_start + row + x0), b = __ldg(cell_start + row + x1 + 1);
        for (unsigned j = a; j < b; j++) {
            if ((int)j == i) continue;                      // the point itself, by index (simple_knn.cu:149,171): duplicates DO count
            const float4 q = __ldg(sorted + j);
            const float dx = q.x - me.x, dy = q.y - me.y, dz = q.z - me.z;
            knn_update3(dx * dx + dy * dy + dz * dz, best);
        }
    };
    for (int r = 0; r <= rmax; r++) {
        // everything not yet visited lies at least r * h_min away (a whole shell of cells lies in between)
        if (r > 0) { const float bound = (float)(r - 1) * g.h_min; if (g.h_min < FLT_MAX && best[2] <= bound * bound) break; }
        const int z0 = max(0, cz - r), z1 = min(g.dim[2] - 1, cz + r), y0 = max(0, cy - r), y1 = min(g.dim[1] - 1, cy + r);
        const int x0 = max(0, cx - r), x1 = min(g.dim[0] - 1, cx + r);
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                const int row = (z * g.dim[1] + y) * g.dim[0];
                if (max(abs(z - cz), abs(y - cy)) == r) scan_run(row, x0, x1);          // a face row of the shell: whole run
                else {                                                                    // interior row: only its two end cells
                    if (cx - r >= 0) scan_run(row, cx - r, cx - r);
                    if (cx + r <= g.dim[0] - 1) scan_run(row, cx + r, cx + r);
                }
            }
    }
    mean_dists[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;            // simple_knn.cu:182
}

}  // namespace dgr
