// dgr_fields.cuh — SURVEY.md §8 row f4: `GaussianModel.extract_fields` (/root/reference/gs_renderer.py:218-294), the dense
// occupancy field the mesh export samples: occ[x][y][z] = sum over Gaussians of opacity * exp(-1/2 d^T Sigma^-1 d) on a
// resolution^3 grid over linspace(-1,1)^3, where a voxel of block B (the grid is cut into num_blocks^3 blocks) only sees
// the Gaussians whose centre lies strictly inside B's voxel bounding box grown by relax_ratio * 2/num_blocks (:262-266).
//
// The reference walks the 4096 blocks in a Python triple loop and materialises [voxels x Gaussians x 3] tensors per
// block.  Here: masked bounding box (block reduce + ordered-int atomics) -> per-Gaussian state (activations, normalised
// centre, inverse covariance: 48 B) -> counting sort by block cell -> one CTA per block that reads only the cell runs its
// grown box can reach, applies the reference's exact strict-inequality test, compacts the survivors in shared memory and
// evaluates them for its voxels from registers.  No host round trip (the reference's `.item()` at :238 stays on the device).
#pragma once
#include "dgr_common.cuh"
#include "dgr_knn.cuh"

namespace dgr {

struct FieldRec { float4 p; float4 i0; float4 i1; };      // {x, y, z, opacity} {inv_a, inv_b, inv_c, inv_d} {inv_e, inv_f, -, -}
static_assert(sizeof(FieldRec) == 48, "FieldRec");

struct FieldsHeader {
    unsigned lo[3], hi[3];        // masked bounding box, ordered-int encoded (knn_enc)
    unsigned n_masked;
    float center[3];
    float scale;
};

struct FieldsLayout {
    size_t off_hdr, off_start, off_fill, off_cid, off_rec, off_sorted, total;
    size_t cells;
    __host__ __device__ FieldsLayout(int P, int nb) {
        cells = (size_t)nb * nb * nb;
        size_t Pn = P > 0 ? (size_t)P : 1, o = 0;
        auto al = [](size_t v) { return (v + 255) / 256 * 256; };
        off_hdr = o;    o = al(o + sizeof(FieldsHeader));
        off_start = o;  o = al(o + (cells + 1) * 4);
        off_fill = o;   o = al(o + cells * 4);
        off_cid = o;    o = al(o + Pn * 4);
        off_rec = o;    o = al(o + Pn * sizeof(FieldRec));
        off_sorted = o; o = al(o + Pn * sizeof(FieldRec));
        total = o;
    }
};

__device__ __forceinline__ float fields_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
fields_bbox_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ opacity_raw, FieldsHeader *__restrict__ hdr) {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    unsigned n = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        if (fields_sigmoid(__ldg(opacity_raw + i)) > 0.005f) {                     // gs_renderer.py:229
            n++;
#pragma unroll
            for (int a = 0; a < 3; a++) { const float v = __ldg(xyz + 3 * (size_t)i + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    }
    if ((threadIdx.x & 31) == 0 && n) {
        atomicAdd(&hdr->n_masked, n);
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicMin(&hdr->lo[a], knn_enc(lo[a])); atomicMax(&hdr->hi[a], knn_enc(hi[a])); }
    }
}

__device__ __forceinline__ int fields_cell(float x, int nb) { return min(nb - 1, max(0, (int)floorf((x + 1.f) * 0.5f * (float)nb))); }

__global__ void __launch_bounds__(256)
fields_prep_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ opacity_raw, const float *__restrict__ scaling_raw,
                   const float *__restrict__ rotation_raw, FieldsHeader *__restrict__ hdr, int nb, FieldRec *__restrict__ rec,
                   unsigned *__restrict__ cid, unsigned *__restrict__ cell_count, float *__restrict__ center_scale) {
    __shared__ float s_c[3], s_scale;
    if (threadIdx.x == 0) {
        float ext = 0.f;
        const bool any = hdr->n_masked > 0;
        for (int a = 0; a < 3; a++) {
            const float mn = any ? knn_dec(hdr->lo[a]) : 0.f, mx = any ? knn_dec(hdr->hi[a]) : 0.f;
            s_c[a] = (mn + mx) / 2.f;                                               // :237
            ext = fmaxf(ext, mx - mn);
        }
        s_scale = any ? (float)(1.8 / (double)ext) : 0.f;                           // :238: python-float division, then a float32 tensor op
        if (blockIdx.x == 0) {
            for (int a = 0; a < 3; a++) { hdr->center[a] = s_c[a]; if (center_scale) center_scale[a] = s_c[a]; }
            hdr->scale = s_scale;
            if (center_scale) center_scale[3] = s_scale;
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float op = fields_sigmoid(__ldg(opacity_raw + i));
    if (!(op > 0.005f)) { cid[i] = 0xffffffffu; return; }
    const float x = (__ldg(xyz + 3 * (size_t)i) - s_c[0]) * s_scale, y = (__ldg(xyz + 3 * (size_t)i + 1) - s_c[1]) * s_scale,
                z = (__ldg(xyz + 3 * (size_t)i + 2) - s_c[2]) * s_scale;                 // :240
    const float sx = expf(__ldg(scaling_raw + 3 * (size_t)i)) * s_scale, sy = expf(__ldg(scaling_raw + 3 * (size_t)i + 1)) * s_scale,
                sz = expf(__ldg(scaling_raw + 3 * (size_t)i + 2)) * s_scale;             // :241
    float4 q = ldg_f4(rotation_raw + 4 * (size_t)i);
    const float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);             // build_rotation normalises, :86-88
    q = make_float4(q.x / qn, q.y / qn, q.z / qn, q.w / qn);
    float R[9]; quat_to_R(q, R);
    // L = R diag(s); Sigma = L L^T, packed xx xy xz yy yz zz (:110-117, :128-132)
    const float L[9] = { R[0] * sx, R[1] * sy, R[2] * sz, R[3] * sx, R[4] * sy, R[5] * sz, R[6] * sx, R[7] * sy, R[8] * sz };
    const float a = L[0] * L[0] + L[1] * L[1] + L[2] * L[2], b = L[0] * L[3] + L[1] * L[4] + L[2] * L[5], c = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
    const float d = L[3] * L[3] + L[4] * L[4] + L[5] * L[5], e = L[3] * L[6] + L[4] * L[7] + L[5] * L[8], f = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
    const float inv_det = 1.f / (a * d * f + 2.f * e * c * b - e * e * a - c * c * d - b * b * f + 1e-24f);      // :71
    FieldRec r;
    r.p = make_float4(x, y, z, op);
    r.i0 = make_float4((d * f - e * e) * inv_det, (e * c - b * f) * inv_det, (e * b - c * d) * inv_det, (a * f - c * c) * inv_det);
    r.i1 = make_float4((b * c - e * a) * inv_det, (a * d - b * b) * inv_det, 0.f, 0.f);
    rec[i] = r;
    const unsigned cell = (unsigned)((fields_cell(z, nb) * nb + fields_cell(y, nb)) * nb + fields_cell(x, nb));
    cid[i] = cell;
    atomicAdd(&cell_count[cell], 1u);
}

// exclusive scan of n counters (n = num_blocks^3 + 1 <= ~262k) by ONE CTA, in place
__global__ void __launch_bounds__(1024) fields_scan_kernel(unsigned *__restrict__ v, int n) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const unsigned x = i < n ? v[i] : 0u;
        unsigned inc = x;
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
        const unsigned excl = s_carry + s_w[threadIdx.x >> 5] + inc - x;
        if (i < n) v[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
fields_scatter_kernel(int P, const unsigned *__restrict__ cid, const unsigned *__restrict__ cell_start, unsigned *__restrict__ cell_fill,
                      const FieldRec *__restrict__ rec, FieldRec *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const unsigned c = cid[i];
    if (c == 0xffffffffu) return;
    sorted[cell_start[c] + atomicAdd(&cell_fill[c], 1u)] = rec[i];
}

__device__ __forceinline__ float fields_linspace(int i, int res) {              // torch.linspace(-1, 1, res)[i]
    const float step = 2.f / (float)(res - 1);
    return i < res / 2 ? __fmaf_rn(step, (float)i, -1.f) : __fmaf_rn(-step, (float)(res - 1 - i), 1.f);      // one rounding, as torch
}

constexpr int kFieldThreads = 256, kFieldChunk = 256;

// One CTA per block of split^3 voxels; every thread owns VPT of them in registers.
template <int VPT>
__global__ void __launch_bounds__(kFieldThreads)
fields_eval_kernel(int res, int nb, int split, float grow, const unsigned *__restrict__ cell_start, const FieldRec *__restrict__ sorted,
                   float *__restrict__ occ) {
    __shared__ FieldRec s_rec[kFieldChunk];
    __shared__ unsigned s_n;
    const int bx = blockIdx.x % nb, by = (blockIdx.x / nb) % nb, bz = blockIdx.x / (nb * nb);      // block (xi, yi, zi) = (bx, by, bz)
    const int V = split * split * split;
    float px[VPT], py[VPT], pz[VPT], val[VPT];
#pragma unroll
    for (int k = 0; k < VPT; k++) {
        const int v = threadIdx.x + k * kFieldThreads;
        const int lx = v / (split * split), ly = (v / split) % split, lz = v % split;
        px[k] = fields_linspace(bx * split + min(lx, split - 1), res);
        py[k] = fields_linspace(by * split + ly, res);
        pz[k] = fields_linspace(bz * split + lz, res);
        val[k] = 0.f;
    }
    // grown voxel bounding box of the block (:262-266) and the cells it can reach
    const float x0 = fields_linspace(bx * split, res) - grow, x1 = fields_linspace(bx * split + split - 1, res) + grow;
    const float y0 = fields_linspace(by * split, res) - grow, y1 = fields_linspace(by * split + split - 1, res) + grow;
    const float z0 = fields_linspace(bz * split, res) - grow, z1 = fields_linspace(bz * split + split - 1, res) + grow;
    const int cx0 = fields_cell(x0, nb), cx1 = fields_cell(x1, nb), cy0 = fields_cell(y0, nb), cy1 = fields_cell(y1, nb);
    const int cz0 = fields_cell(z0, nb), cz1 = fields_cell(z1, nb);
    for (int cz = cz0; cz <= cz1; cz++)
        for (int cy = cy0; cy <= cy1; cy++) {
            const int row = (cz * nb + cy) * nb;
            const unsigned a = __ldg(cell_start + row + cx0), b = __ldg(cell_start + row + cx1 + 1);      // one contiguous run
            for (unsigned base = a; base < b; base += kFieldChunk) {
                if (threadIdx.x == 0) s_n = 0;
                __syncthreads();
                const unsigned j = base + threadIdx.x;
                if (j < b) {
                    const float4 p = __ldg(&sorted[j].p);
                    if (p.x < x1 && p.x > x0 && p.y < y1 && p.y > y0 && p.z < z1 && p.z > z0) {         // strict, as the reference
                        const unsigned slot = atomicAdd(&s_n, 1u);
                        s_rec[slot].p = p; s_rec[slot].i0 = __ldg(&sorted[j].i0); s_rec[slot].i1 = __ldg(&sorted[j].i1);
                    }
                }
                __syncthreads();
                const unsigned n = s_n;
                for (unsigned g = 0; g < n; g++) {
                    const float4 p = s_rec[g].p, i0 = s_rec[g].i0, i1 = s_rec[g].i1;
#pragma unroll
                    for (int k = 0; k < VPT; k++) {
                        const float dx = px[k] - p.x, dy = py[k] - p.y, dz = pz[k] - p.z;
                        // gaussian_3d_coeff, :78
                        const float power = -0.5f * (dx * dx * i0.x + dy * dy * i0.w + dz * dz * i1.y) - dx * dy * i0.y - dx * dz * i0.z - dy * dz * i1.x;
                        if (power <= 0.f) val[k] += p.w * ex2_approx(power * kLog2e);                   // power > 0 -> exp(-1e10) = 0 (:81)
                    }
                }
                __syncthreads();
            }
        }
#pragma unroll
    for (int k = 0; k < VPT; k++) {
        const int v = threadIdx.x + k * kFieldThreads;
        if (v < V) {
            const int lx = v / (split * split), ly = (v / split) % split, lz = v % split;
            occ[((size_t)(bx * split + lx) * res + (by * split + ly)) * res + (bz * split + lz)] = val[k];
        }
    }
}

}  // namespace dgr
