// dgr_api.cu — C ABI of libdgr_b200.so (see include/dgr_b200.h).  Host-side glue only: argument checks, scratch
// layout, kernel launches on the caller's stream.  No torch, no CPU fallback: without a CUDA device every compute
// entry point fails with an error string.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dgr_b200.h"
#include "dgr_backward.cuh"
#include "dgr_binning.cuh"
#include "dgr_common.cuh"
#include "dgr_preprocess.cuh"
#include "dgr_render.cuh"

using namespace dgr;

namespace {
thread_local std::string g_err;
thread_local uint64_t g_launches = 0;

int fail(int code, const char *what, const char *detail = nullptr) {
    g_err = what;
    if (detail) { g_err += ": "; g_err += detail; }
    return code;
}
#define DGR_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail((int)e_, #call, cudaGetErrorString(e_));            \
    } while (0)
// optional per-kernel CUDA-event timing (bench.py's live roofline): events bracket every launch on its stream
struct ProfRec { const char *name; cudaEvent_t a, b; };
thread_local bool g_prof_on = false;
thread_local std::vector<ProfRec> g_prof;
inline void prof_begin(const char *name, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
inline void prof_end(cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEventRecord(g_prof.back().b, st);
}
#define DGR_KERNEL(name, st, dbg, ...)          \
    do {                                         \
        prof_begin(name, st);                    \
        __VA_ARGS__;                             \
        prof_end(st);                            \
        DGR_LAUNCHED(st, dbg);                   \
    } while (0)

#define DGR_LAUNCHED(s, dbg)                                                                   \
    do {                                                                                       \
        g_launches++;                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) return fail((int)e_, "kernel launch", cudaGetErrorString(e_));  \
        if (dbg) { e_ = cudaStreamSynchronize(s);                                              \
            if (e_ != cudaSuccess) return fail((int)e_, "kernel execution", cudaGetErrorString(e_)); } \
    } while (0)

int check_settings(const DgrSettings *s) {
    if (!s) return fail(-1, "settings is NULL");
    if (s->image_height <= 0 || s->image_width <= 0) return fail(-1, "image size must be positive");
    if (s->image_height > 65535 || s->image_width > 65535) return fail(-1, "image size above 65535 not supported");
    if (s->sh_degree < 0 || s->sh_degree > 3) return fail(-1, "sh_degree must be 0..3");
    if (!s->bg || !s->viewmatrix || !s->projmatrix) return fail(-1, "bg / viewmatrix / projmatrix must not be NULL");
    return 0;
}
int check_gaussians(const DgrSettings *s, const DgrGaussians *g) {
    if (!g) return fail(-1, "gaussians is NULL");
    if (g->P < 0) return fail(-1, "P < 0");
    if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
        return fail(-2, "Please provide excatly one of either SHs or precomputed colors!");
    const bool sr = g->scales != nullptr && g->rotations != nullptr;
    if (((g->scales == nullptr || g->rotations == nullptr) && g->cov3D_precomp == nullptr) ||
        ((g->scales != nullptr || g->rotations != nullptr) && g->cov3D_precomp != nullptr))
        return fail(-2, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    (void)sr;
    if (g->P > 0 && (!g->means3D || !g->opacities)) return fail(-1, "means3D / opacities must not be NULL");
    if (g->shs) {
        if (!s->campos) return fail(-1, "campos must not be NULL with SHs");
        const int nb = (s->sh_degree + 1) * (s->sh_degree + 1);
        if (g->M < nb) return fail(-1, "shs has fewer coefficients than sh_degree needs");
        if (g->M > 16) return fail(-1, "shs with more than 16 coefficients per channel not supported");
    }
    return 0;
}

size_t sort_temp_bytes(uint64_t cap) {
    size_t bytes = 0;
    unsigned long long *k = nullptr; unsigned *v = nullptr;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, k, v, v, (int)(cap > 0 ? cap : 1), 0, 64);
    if (e != cudaSuccess) { cudaGetLastError(); bytes = (size_t)(cap > 0 ? cap : 1) * 16 + (1u << 20); }   // conservative fallback
    return bytes;
}

size_t geom_total(int P) { GeomLayout L(P); return L.total + align_up((size_t)(P > 0 ? P : 1) * kGradRecFloats * 4, 256); }

template <int DEG, bool HAS_SH, bool HAS_COV>
void launch_pre_fwd(const DgrSettings *s, const DgrGaussians *g, int *radii, char *geom, const GeomLayout &L, cudaStream_t st) {
    preprocess_fwd_kernel<DEG, HAS_SH, HAS_COV><<<L.nblocks, kPreThreads, 0, st>>>(
        g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
        s->campos, g->means3D, g->shs, g->colors_precomp, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii,
        reinterpret_cast<GeomHeader *>(geom), reinterpret_cast<unsigned long long *>(geom + L.off_status),
        reinterpret_cast<Rec *>(geom + L.off_rec), reinterpret_cast<unsigned *>(geom + L.off_offsets),
        reinterpret_cast<unsigned *>(geom + L.off_touched), L.nblocks);
}

template <int DEG, bool HAS_SH, bool HAS_COV>
void launch_pre_bwd(const DgrSettings *s, const DgrGaussians *g, const int *radii, const float *grad_rec,
                    const DgrGaussianGrads *o, cudaStream_t st) {
    const int nb = (g->P + kPreThreads - 1) / kPreThreads;
    preprocess_bwd_kernel<DEG, HAS_SH, HAS_COV><<<nb, kPreThreads, 0, st>>>(
        g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
        s->campos, g->means3D, g->shs, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii, grad_rec,
        o->dL_dmeans3D, o->dL_dmeans2D, o->dL_dshs, o->dL_dcolors_precomp, o->dL_dopacities, o->dL_dscales, o->dL_drotations,
        o->dL_dcov3D_precomp, o->accumulate);
}

__global__ void debug_geom_kernel(int P, const Rec *rec, const unsigned *touched, float *mean_px, float *depth, float *conic,
                                  float *rgb, int *aabb, unsigned *tiles_touched) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    const Rec r = rec[g];
    if (mean_px) { mean_px[2 * g] = r.q0.x; mean_px[2 * g + 1] = r.q0.y; }
    if (depth) depth[g] = r.q1.z;
    if (conic) { conic[3 * g] = r.q0.z * (-2.f / kLog2e); conic[3 * g + 1] = r.q0.w * (-1.f / kLog2e); conic[3 * g + 2] = r.q1.x * (-2.f / kLog2e); }
    if (rgb) { rgb[3 * g] = r.q2.x; rgb[3 * g + 1] = r.q2.y; rgb[3 * g + 2] = r.q2.z; }
    if (aabb) {
        const unsigned ax = __float_as_uint(r.q1.w), ay = __float_as_uint(r.q2.w);
        aabb[4 * g] = (int)(ax & 0xffffu); aabb[4 * g + 1] = (int)(ay & 0xffffu); aabb[4 * g + 2] = (int)(ax >> 16); aabb[4 * g + 3] = (int)(ay >> 16);
    }
    if (tiles_touched) tiles_touched[g] = touched[g];
}
}  // namespace

extern "C" {

int dgr_abi_version(void) { return DGR_ABI_VERSION; }

void dgr_profile_enable(int on) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
}
// Synchronises the recorded events; writes up to `max` (name, milliseconds) pairs, names as a '\n'-joined string.
int dgr_profile_collect(char *names, size_t names_bytes, float *ms, int max) {
    int n = 0; size_t off = 0;
    if (names && names_bytes) names[0] = 0;
    for (auto &r : g_prof) {
        if (n >= max) break;
        if (cudaEventSynchronize(r.b) != cudaSuccess) break;
        float t = 0.f; cudaEventElapsedTime(&t, r.a, r.b);
        ms[n] = t;
        size_t L = strlen(r.name);
        if (names && off + L + 2 < names_bytes) { memcpy(names + off, r.name, L); off += L; names[off++] = '\n'; names[off] = 0; }
        n++;
    }
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    return n;
}
const char *dgr_last_error(void) { return g_err.c_str(); }
uint64_t dgr_launch_count(void) { return g_launches; }
void dgr_reset_launch_count(void) { g_launches = 0; }

size_t dgr_geom_bytes(int32_t P) { return geom_total(P); }
size_t dgr_image_bytes(int32_t H, int32_t W) { ImageLayout L(H, W); return L.total + align_up((size_t)H * W * 4, 256); }
size_t dgr_binning_bytes(uint64_t cap, int32_t H, int32_t W) {
    (void)H; (void)W;
    BinningLayout L(cap, sort_temp_bytes(cap));
    return L.total;
}

int dgr_forward_preprocess(const DgrSettings *s, const DgrGaussians *g, void *geom_v, int32_t *radii,
                           uint64_t *n_instances_host, void *stream) {
    if (int e = check_settings(s)) return e;
    if (int e = check_gaussians(s, g)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v;
    if (!geom) return fail(-1, "geom scratch is NULL");
    GeomLayout L(g->P);
    DGR_CUDA(cudaMemsetAsync(geom, 0, L.off_rec, st));
    if (g->P > 0) {
        if (!radii) return fail(-1, "radii is NULL");
        const bool sh = g->shs != nullptr, cov = g->cov3D_precomp != nullptr;
        prof_begin("preprocess_fwd", st);
        if (sh) {
#define DGR_DISPATCH_DEG(D)                                                     \
    case D:                                                                     \
        if (cov) launch_pre_fwd<D, true, true>(s, g, radii, geom, L, st);       \
        else launch_pre_fwd<D, true, false>(s, g, radii, geom, L, st);          \
        break;
            switch (s->sh_degree) { DGR_DISPATCH_DEG(0) DGR_DISPATCH_DEG(1) DGR_DISPATCH_DEG(2) DGR_DISPATCH_DEG(3) }
#undef DGR_DISPATCH_DEG
        } else {
            if (cov) launch_pre_fwd<0, false, true>(s, g, radii, geom, L, st);
            else launch_pre_fwd<0, false, false>(s, g, radii, geom, L, st);
        }
        prof_end(st);
        DGR_LAUNCHED(st, s->debug);
    }
    if (n_instances_host)
        DGR_CUDA(cudaMemcpyAsync(n_instances_host, geom, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    return 0;
}

int dgr_forward_render(const DgrSettings *s, const DgrGaussians *g, void *geom_v, void *binning_v,
                       uint64_t capacity, void *image_v, const DgrImages *out, void *stream) {
    if (int e = check_settings(s)) return e;
    if (!g || !geom_v || !image_v || !out) return fail(-1, "NULL argument");
    if (!out->color || !out->depth || !out->alpha) return fail(-1, "output images must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v, *binning = (char *)binning_v, *image = (char *)image_v;
    const int H = s->image_height, W = s->image_width;
    GeomLayout GL(g->P);
    ImageLayout IL(H, W);
    const size_t tiles = (size_t)IL.gx * IL.gy;
    uint2 *ranges = reinterpret_cast<uint2 *>(image + IL.off_ranges);
    unsigned *n_contrib = reinterpret_cast<unsigned *>(image + IL.off_ncontrib);
    DGR_CUDA(cudaMemsetAsync(ranges, 0, tiles * sizeof(uint2), st));

    // v1: the instance count is read back (one host sync per forward, as the reference op does).
    unsigned long long n_inst = 0;
    DGR_CUDA(cudaMemcpyAsync(&n_inst, geom, sizeof(n_inst), cudaMemcpyDeviceToHost, st));
    DGR_CUDA(cudaStreamSynchronize(st));
    if (n_inst >= 0xffffffffull) return fail(-3, "more than 2^32-1 tile instances");
    unsigned long long n = n_inst < capacity ? n_inst : capacity;
    const Rec *rec_sorted = nullptr;
    if (n > 0) {
        if (!binning) return fail(-1, "binning scratch is NULL");
        const size_t temp_bytes = sort_temp_bytes(capacity);
        BinningLayout BL(capacity, temp_bytes);
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(binning + BL.off_keys);
        unsigned long long *keys_alt = reinterpret_cast<unsigned long long *>(binning + BL.off_keys_alt);
        unsigned *vals = reinterpret_cast<unsigned *>(binning + BL.off_vals);
        unsigned *vals_alt = reinterpret_cast<unsigned *>(binning + BL.off_vals_alt);
        Rec *recs = reinterpret_cast<Rec *>(binning + BL.off_rec);
        const Rec *rec = reinterpret_cast<const Rec *>(geom + GL.off_rec);
        prof_begin("emit_instances", st);
        emit_instances_kernel<<<(g->P + 255) / 256, 256, 0, st>>>(
            g->P, IL.gx, rec, reinterpret_cast<const unsigned *>(geom + GL.off_offsets),
            reinterpret_cast<const unsigned *>(geom + GL.off_touched), capacity, keys, vals);
        prof_end(st);
        DGR_LAUNCHED(st, s->debug);
        int tile_bits = 0;
        while (((size_t)1 << tile_bits) < tiles) tile_bits++;
        size_t tb = temp_bytes;
        prof_begin("radix_sort(cub)", st);
        DGR_CUDA(cub::DeviceRadixSort::SortPairs(binning + BL.off_temp, tb, keys, keys_alt, vals, vals_alt, (int)n, 0, 32 + tile_bits, st));
        prof_end(st);
        g_launches += 1;   // counted as one library step (CUB launches several kernels internally)
        prof_begin("ranges_gather", st);
        ranges_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, keys_alt, vals_alt, rec, ranges, recs);
        prof_end(st);
        DGR_LAUNCHED(st, s->debug);
        rec_sorted = recs;
    }
    float *final_T = reinterpret_cast<float *>(image + IL.total);
    prof_begin("render_fwd", st);
    render_fwd_kernel<<<(unsigned)tiles, kTileThreads, 0, st>>>(H, W, IL.gx, ranges, rec_sorted, s->bg, out->color, out->depth,
                                                               out->alpha, n_contrib, final_T);
    prof_end(st);
    DGR_LAUNCHED(st, s->debug);
    return 0;
}

int dgr_backward(const DgrSettings *s, const DgrGaussians *g, void *geom_v, const void *binning_v,
                 uint64_t capacity, const void *image_v, const int32_t *radii, const float *out_alpha, const DgrImageGrads *gin,
                 const DgrGaussianGrads *gout, void *stream) {
    (void)out_alpha;
    if (int e = check_settings(s)) return e;
    if (int e = check_gaussians(s, g)) return e;
    if (!geom_v || !image_v || !gin || !gout) return fail(-1, "NULL argument");
    if (g->P == 0) return 0;
    if (!radii) return fail(-1, "radii is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v;
    const char *binning = (const char *)binning_v, *image = (const char *)image_v;
    const int H = s->image_height, W = s->image_width;
    GeomLayout GL(g->P);
    ImageLayout IL(H, W);
    const size_t tiles = (size_t)IL.gx * IL.gy;
    float *grad_rec = reinterpret_cast<float *>(geom + GL.total);
    DGR_CUDA(cudaMemsetAsync(grad_rec, 0, (size_t)g->P * kGradRecFloats * 4, st));
    if (binning && capacity > 0) {
        const uint64_t cap = capacity;
        BinningLayout BL(cap, sort_temp_bytes(cap));
        const unsigned *ids_sorted = reinterpret_cast<const unsigned *>(binning + BL.off_vals_alt);
        const Rec *recs = reinterpret_cast<const Rec *>(binning + BL.off_rec);
        const float *final_T = reinterpret_cast<const float *>(image + IL.total);
        prof_begin("render_bwd", st);
        render_bwd_kernel<<<(unsigned)tiles, kTileThreads, 0, st>>>(
            H, W, IL.gx, reinterpret_cast<const uint2 *>(image + IL.off_ranges), recs, ids_sorted, s->bg, final_T,
            reinterpret_cast<const unsigned *>(image + IL.off_ncontrib), gin->dL_dcolor, gin->dL_ddepth, gin->dL_dalpha, grad_rec);
        prof_end(st);
        DGR_LAUNCHED(st, s->debug);
    }
    const bool sh = g->shs != nullptr, cov = g->cov3D_precomp != nullptr;
    prof_begin("preprocess_bwd", st);
    if (sh) {
#define DGR_DISPATCH_DEG(D)                                                       \
    case D:                                                                       \
        if (cov) launch_pre_bwd<D, true, true>(s, g, radii, grad_rec, gout, st);  \
        else launch_pre_bwd<D, true, false>(s, g, radii, grad_rec, gout, st);     \
        break;
        switch (s->sh_degree) { DGR_DISPATCH_DEG(0) DGR_DISPATCH_DEG(1) DGR_DISPATCH_DEG(2) DGR_DISPATCH_DEG(3) }
#undef DGR_DISPATCH_DEG
    } else {
        if (cov) launch_pre_bwd<0, false, true>(s, g, radii, grad_rec, gout, st);
        else launch_pre_bwd<0, false, false>(s, g, radii, grad_rec, gout, st);
    }
    prof_end(st);
    DGR_LAUNCHED(st, s->debug);
    return 0;
}

int dgr_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                     uint8_t *present, void *stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(-1, "bad argument");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, viewmatrix, present);
    DGR_LAUNCHED(st, 0);
    return 0;
}

int dgr_debug_geom(int32_t P, const void *geom_v, float *mean_px, float *depth, float *conic, float *rgb,
                   int32_t *aabb, uint32_t *tiles_touched, void *stream) {
    if (P <= 0) return 0;
    if (!geom_v) return fail(-1, "geom is NULL");
    const char *geom = (const char *)geom_v;
    GeomLayout L(P);
    cudaStream_t st = (cudaStream_t)stream;
    debug_geom_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, reinterpret_cast<const Rec *>(geom + L.off_rec),
                                                     reinterpret_cast<const unsigned *>(geom + L.off_touched), mean_px, depth,
                                                     conic, rgb, aabb, tiles_touched);
    DGR_LAUNCHED(st, 0);
    return 0;
}

}  // extern "C"
