// dgr_api.cu — C ABI of libdgr_b200.so (see include/dgr_b200.h).  Host-side glue only: argument checks, scratch
// layout, kernel launches on the caller's stream.  No torch, no library kernels (no CUB/cuBLAS), no CPU fallback:
// without a CUDA device every compute entry point fails with an error string.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dgr_b200.h"
#include "dgr_backward.cuh"
#include "dgr_binning.cuh"
#include "dgr_collective.cuh"
#include "dgr_densify.cuh"
#include "dgr_knn.cuh"
#include "dgr_fields.cuh"
#include "dgr_optim.cuh"
#include "dgr_common.cuh"
#include "dgr_preprocess.cuh"
#include "dgr_render.cuh"

using namespace dgr;

namespace {
thread_local std::string g_err;
thread_local uint64_t g_launches = 0;

// NVTX range around every C-ABI entry point (tracing): header-only NVTX v3, a no-op unless a tool is attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int fail(int code, const char *what, const char *detail = nullptr) {
    g_err = what;
    if (detail) { g_err += ": "; g_err += detail; }
    return code;
}

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

#define DGR_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail((int)e_, #call, cudaGetErrorString(e_));            \
    } while (0)
// launch a kernel (the statement in __VA_ARGS__), count it, time it when profiling, check it
#define DGR_KERNEL(name, st, dbg, ...)                                                         \
    do {                                                                                       \
        prof_begin(name, st);                                                                  \
        __VA_ARGS__;                                                                           \
        prof_end(st);                                                                          \
        g_launches++;                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) return fail((int)e_, name, cudaGetErrorString(e_));             \
        if (dbg) { e_ = cudaStreamSynchronize(st);                                             \
            if (e_ != cudaSuccess) return fail((int)e_, name, cudaGetErrorString(e_)); }       \
    } while (0)

// Kernel launch with the programmatic-dependent-launch attribute: the kernel may become resident while its predecessor in
// the stream (which must be a kernel, not a copy / memset / event) is still running; it calls pdl_wait() before it touches
// anything the predecessor produces.  g_pdl = 0 (DGR_PDL=0 or dgr_set_tuning bit 3) launches the same kernels the ordinary way.
std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load();
    if (v < 0) { const char *e = getenv("DGR_PDL"); v = (e && e[0] == '0') ? 0 : 1; g_pdl = v; }
    return v != 0;
}
template <class... KArgs, class... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const bool on = pdl && pdl_enabled();
    cfg.attrs = on ? attr : nullptr; cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);       // errors surface through cudaGetLastError in DGR_KERNEL
}

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
    if (g->P == 0) return 0;                       // empty cloud: empty tensors have NULL data pointers
    if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
        return fail(-2, "Please provide excatly one of either SHs or precomputed colors!");
    if (((g->scales == nullptr || g->rotations == nullptr) && g->cov3D_precomp == nullptr) ||
        ((g->scales != nullptr || g->rotations != nullptr) && g->cov3D_precomp != nullptr))
        return fail(-2, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    if (g->P > 0 && (!g->means3D || !g->opacities)) return fail(-1, "means3D / opacities must not be NULL");
    if (g->activations) {
        if (!g->shs || g->cov3D_precomp || !g->scales || !g->rotations)
            return fail(-2, "activations: raw parameters need features_dc (shs), scaling and rotation; no cov3D_precomp");
        if (g->M > 1 && !g->shs_rest) return fail(-2, "activations: shs_rest (_features_rest) is NULL but M > 1");
    }
    if (g->shs) {
        if (!s->campos) return fail(-1, "campos must not be NULL with SHs");
        const int nb = (s->sh_degree + 1) * (s->sh_degree + 1);
        if (g->M < nb) return fail(-1, "shs has fewer coefficients than sh_degree needs");
        if (g->M > 16) return fail(-1, "shs with more than 16 coefficients per channel not supported");
    }
    return 0;
}

constexpr size_t kMaxTileSmem = 200 * 1024;      // tile-histogram rows live in shared memory: up to 51200 tiles

// tuning knobs of dgr_set_tuning bits 24-26 (atomics, like the ones further down)
std::atomic<int> g_stage_fwd{1}, g_stage_bwd{1};   // per-Gaussian kernels: inputs staged through shared memory by bulk TMA (dgr_preprocess.cuh)
std::atomic<int> g_sort_fine{0};               // per-tile sort with 2048 instead of 1024 depth buckets (A/B, bit 28)
std::atomic<int> g_rowsum{1};                  // backward render, step 2: per-row form of the dy moments (dgr_render.cuh, ROWSUM)

// Can the per-Gaussian kernels stage their inputs with bulk TMA (dgr_preprocess.cuh)?  SH + scale / rotation inputs, every base
// pointer 16-byte aligned (cp.async.bulk), P at least one full warp.
inline bool stage_inputs_ok(const DgrGaussians *g, bool raw) {
    if (!g->shs || g->cov3D_precomp || !g->scales || !g->rotations || g->P < 32) return false;
    uintptr_t bits = (uintptr_t)g->means3D | (uintptr_t)g->scales | (uintptr_t)g->rotations | (uintptr_t)g->opacities | (uintptr_t)g->shs;
    if (raw && g->M > 1) bits |= (uintptr_t)g->shs_rest;
    return (bits & 15u) == 0;
}
constexpr size_t kMaxKernelSmem = 227 * 1024;

template <int DEG, bool HAS_SH, bool HAS_COV, bool RAW>
void launch_pre_fwd(const DgrSettings *s, const DgrGaussians *g, int *radii, char *geom, const GeomLayout &L, unsigned *tile_count,
                    cudaStream_t st) {
    unsigned *run_matrix = reinterpret_cast<unsigned *>(geom + L.off_runs);
    const size_t hist = (size_t)L.tiles * 4;
    if constexpr (HAS_SH && !HAS_COV) {
        // inputs staged through shared memory by bulk TMA: [tile histogram | 8 warps x 32 x (44 + 12 M) bytes]
        const size_t stage_off = align_up(hist, 128);
        const size_t smem = stage_off + (size_t)(kPreThreads / 32) * stage_warp_floats(g->M) * 4;
        if (g_stage_fwd.load() && stage_inputs_ok(g, RAW) && smem + 1024 <= kMaxKernelSmem) {
            if (smem > 48 * 1024)
                cudaFuncSetAttribute(preprocess_fwd_kernel<DEG, HAS_SH, HAS_COV, RAW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            launch_k(preprocess_fwd_kernel<DEG, HAS_SH, HAS_COV, RAW, true>, dim3(L.nblocks), dim3(kPreThreads), smem, st, false,
                g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
                s->campos, g->means3D, g->shs, g->shs_rest, g->colors_precomp, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii,
                reinterpret_cast<Rec *>(geom + L.off_rec), reinterpret_cast<unsigned *>(geom + L.off_touched),
                tile_count, run_matrix, L.tiles, L.iters, (int)stage_off);
            return;
        }
    }
    const size_t smem = hist;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(preprocess_fwd_kernel<DEG, HAS_SH, HAS_COV, RAW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(preprocess_fwd_kernel<DEG, HAS_SH, HAS_COV, RAW, false>, dim3(L.nblocks), dim3(kPreThreads), smem, st, false,
        g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
        s->campos, g->means3D, g->shs, g->shs_rest, g->colors_precomp, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii,
        reinterpret_cast<Rec *>(geom + L.off_rec), reinterpret_cast<unsigned *>(geom + L.off_touched),
        tile_count, run_matrix, L.tiles, L.iters, 0);
}

template <int DEG, bool HAS_SH, bool HAS_COV, bool RAW>
void launch_pre_bwd(const DgrSettings *s, const DgrGaussians *g, const int *radii, const unsigned *touched, const float *grad_rec,
                    const DgrGaussianGrads *o, cudaStream_t st) {
    const int nb = (g->P + kPreThreads - 1) / kPreThreads;
    PeerPush push;
    memset(&push, 0, sizeof(push));
    if (o->push) {
        push.per = (int)o->push->gaussians_per_owner;
        for (int w = 0; w < o->push->world && w < kMaxPeers; w++) push.delta[w] = o->push->delta_floats[w];
    }
    if constexpr (HAS_SH && !HAS_COV) {
        // inputs staged by bulk TMA; the same per-warp slice of dynamic shared memory is the warp's output image afterwards
        const int wfl = max(stage_warp_floats(g->M), kStageFloats);
        const size_t smem = (size_t)(kPreThreads / 32) * wfl * 4;
        if (g_stage_bwd.load() && stage_inputs_ok(g, RAW) && smem + 1024 <= kMaxKernelSmem) {
            if (smem > 48 * 1024)
                cudaFuncSetAttribute(preprocess_bwd_kernel<DEG, HAS_SH, HAS_COV, RAW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            launch_k(preprocess_bwd_kernel<DEG, HAS_SH, HAS_COV, RAW, true>, dim3(nb), dim3(kPreThreads), smem, st, true,
                g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
                s->campos, g->means3D, g->shs, g->shs_rest, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii, touched, grad_rec,
                o->dL_dmeans3D, o->dL_dmeans2D, o->dL_dshs, o->dL_dcolors_precomp, o->dL_dopacities, o->dL_dscales, o->dL_drotations,
                o->dL_dcov3D_precomp, o->dL_dshs_rest, o->xyz_gradient_accum, o->denom, o->max_radii2D, o->accumulate, push, wfl);
            return;
        }
    }
    launch_k(preprocess_bwd_kernel<DEG, HAS_SH, HAS_COV, RAW, false>, dim3(nb), dim3(kPreThreads), 0, st, true,
        g->P, g->M, s->image_height, s->image_width, s->tanfovx, s->tanfovy, s->scale_modifier, s->viewmatrix, s->projmatrix,
        s->campos, g->means3D, g->shs, g->shs_rest, g->opacities, g->scales, g->rotations, g->cov3D_precomp, radii, touched, grad_rec,
        o->dL_dmeans3D, o->dL_dmeans2D, o->dL_dshs, o->dL_dcolors_precomp, o->dL_dopacities, o->dL_dscales, o->dL_drotations,
        o->dL_dcov3D_precomp, o->dL_dshs_rest, o->xyz_gradient_accum, o->denom, o->max_radii2D, o->accumulate, push, 0);
}

#define DGR_DISPATCH(FN, ...)                                                                              \
    do {                                                                                                   \
        const bool sh_ = g->shs != nullptr, cov_ = g->cov3D_precomp != nullptr;                            \
        if (g->activations) switch (s->sh_degree) {       /* raw parameters: SH + scale/rotation only */     \
            case 0: FN<0, true, false, true>(__VA_ARGS__); break;                                           \
            case 1: FN<1, true, false, true>(__VA_ARGS__); break;                                           \
            case 2: FN<2, true, false, true>(__VA_ARGS__); break;                                           \
            default: FN<3, true, false, true>(__VA_ARGS__); break;                                          \
        }                                                                                                  \
        else if (!sh_) { if (cov_) FN<0, false, true, false>(__VA_ARGS__); else FN<0, false, false, false>(__VA_ARGS__); } \
        else switch (s->sh_degree) {                                                                       \
            case 0: if (cov_) FN<0, true, true, false>(__VA_ARGS__); else FN<0, true, false, false>(__VA_ARGS__); break;  \
            case 1: if (cov_) FN<1, true, true, false>(__VA_ARGS__); else FN<1, true, false, false>(__VA_ARGS__); break;  \
            case 2: if (cov_) FN<2, true, true, false>(__VA_ARGS__); else FN<2, true, false, false>(__VA_ARGS__); break;  \
            default: if (cov_) FN<3, true, true, false>(__VA_ARGS__); else FN<3, true, false, false>(__VA_ARGS__); break; \
        }                                                                                                  \
    } while (0)

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
    if (tiles_touched) tiles_touched[g] = touched[g] & 0x1fffffffu;
}

// tuning knobs (dgr_set_tuning): pixels per lane of the render kernels, heaviest-first tile order on/off.  Atomics: any host
// thread may change them while another launches (each launch reads every knob once).
std::atomic<int> g_ppl_fwd{1}, g_ppl_bwd{1}, g_u_fwd{0}, g_u_bwd{0};      // u = 0: the default batching of that sub-tile shape
std::atomic<bool> g_no_order{false};
std::atomic<int> g_two_ended{0};
std::atomic<int> g_cost_order{1};
std::atomic<int> g_lazy{0};            // record staging of the render kernels: 0 = by id when the sorted copy would exceed 2x L2, 1 = always sorted copy, 2 = always by id      // backward work ordered by the cost the forward measured (dgr_render.cuh)       // work queue of the render kernels consumed from both ends (dgr_render.cuh)
std::atomic<int> g_cta_fwd{0}, g_cta_bwd{0};   // persistent CTAs per SM of the render kernels (0 = as many as fit)

using SmS = SortSmem<kSortSmallThreads, kSortSmallCap, kSortSmallBuckets>;
using SmSF = SortSmem<kSortSmallThreads, kSortSmallCap, kSortSmallBucketsFine>;
using SmB = SortSmem<kSortBigThreads, kSortBigCap, kSortBigBuckets>;

// Per-DEVICE launch state: cudaFuncSetAttribute (dynamic shared memory above 48 KB) and the occupancy-derived persistent
// grids are properties of a (function, device) pair, so they are kept per device ordinal and set up once under a mutex.
struct DevInfo {
    bool ready = false;
    int sms = 148;
    size_t l2_bytes = 126u << 20;
    int big_grid = 148;               // big-tile sorter: one CTA per SM
    std::map<const void *, int> grids;   // persistent (occupancy x SMs) grid of every render kernel instantiation used so far
};
constexpr int kMaxDevices = 64;
DevInfo g_dev[kMaxDevices];
std::mutex g_dev_mu;

DevInfo *dev_info() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { fail(-1, "cudaGetDevice failed or device ordinal above 63"); return nullptr; }
    std::lock_guard<std::mutex> lock(g_dev_mu);
    DevInfo &d = g_dev[dev];
    if (d.ready) return &d;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) d.sms = n;
    d.big_grid = d.sms;
    int l2 = 0;
    if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) == cudaSuccess && l2 > 0) d.l2_bytes = (size_t)l2;
    cudaError_t e = cudaFuncSetAttribute(emit_instances_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sort_gather_kernel<kSortSmallBuckets>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmS::bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sort_gather_kernel<kSortSmallBucketsFine>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmSF::bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sort_gather_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmB::bytes);
    if (e != cudaSuccess) { fail((int)e, "cudaFuncSetAttribute", cudaGetErrorString(e)); return nullptr; }
    d.ready = true;
    return &d;
}

// persistent grid of a render kernel on the current device: (CTAs that fit on one SM) x SMs, computed once per (device, kernel)
template <class K>
int persistent_grid(DevInfo *d, K kernel, int threads, size_t smem) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    const void *key = reinterpret_cast<const void *>(kernel);
    auto it = d->grids.find(key);
    if (it != d->grids.end()) return it->second;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int g = per_sm * d->sms;
    d->grids[key] = g;
    return g;
}
}  // namespace

extern "C" {

int dgr_abi_version(void) { return DGR_ABI_VERSION; }
const char *dgr_last_error(void) { return g_err.c_str(); }
uint64_t dgr_launch_count(void) { return g_launches; }
void dgr_reset_launch_count(void) { g_launches = 0; }

void dgr_profile_enable(int on) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
}
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

int dgr_set_tuning(int ppl_fwd, int ppl_bwd, int tile_order) {
    if ((ppl_fwd != 1 && ppl_fwd != 2 && ppl_fwd != 4) || (ppl_bwd != 1 && ppl_bwd != 2 && ppl_bwd != 4)) return fail(-1, "ppl must be 1, 2 or 4");
    g_ppl_fwd = ppl_fwd; g_ppl_bwd = ppl_bwd == 4 ? 2 : ppl_bwd;          // the backward has 8x4 and 8x8 sub-tiles
    g_no_order = (tile_order & 3) == 0;
    g_pdl = (tile_order & 8) ? 0 : 1;                   // bit 3: programmatic dependent launches OFF (A/B switch)
    g_u_fwd = (tile_order >> 4) & 7;                    // bits 4-6 / 8-10: hits evaluated together by the forward / backward
    g_u_bwd = (tile_order >> 8) & 7;                    // render kernels (1, 2 or 4; 0 = default of the sub-tile shape)
    g_cta_fwd = (tile_order >> 12) & 15;                // bits 12-15 / 16-19: persistent CTAs per SM of the forward / backward
    g_cta_bwd = (tile_order >> 16) & 15;                // render kernels (0 = as many as fit)
    g_two_ended = (tile_order >> 20) & 1;               // bit 20: work queue consumed from both ends (experiment; measured slower)
    g_cost_order = (tile_order >> 21) & 1 ? 0 : 1;      // bit 21: backward work in tile-population order instead of measured cost
    g_lazy = (tile_order >> 22) & 3;                    // bits 22-23: record staging 0 = auto, 1 = always the sorted copy, 2 = always by id
    g_stage_fwd = (tile_order >> 24) & 1 ? 0 : 1;       // bit 24 / 25: per-Gaussian forward / backward kernel reads its inputs with
    g_stage_bwd = (tile_order >> 25) & 1 ? 0 : 1;       // per-thread global loads instead of bulk-TMA staging (A/B switch)
    g_sort_fine = (tile_order >> 28) & 1;               // bit 28: per-tile sort with 2048 depth buckets (A/B switch)
    g_rowsum = (tile_order >> 26) & 1 ? 0 : 1;          // bit 26: backward render step 2 in the per-pixel form (A/B switch; <1,2> non-lazy only)
    return 0;
}

void *dgr_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { fail(-1, "cudaEventCreate failed"); return nullptr; }
    return (void *)e;
}
int dgr_event_synchronize(void *ev) { DGR_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); return 0; }
void dgr_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }

size_t dgr_geom_bytes(int32_t P, int32_t H, int32_t W) { return GeomLayout(P, H, W).total; }
size_t dgr_image_bytes(int32_t H, int32_t W) { return ImageLayout(H, W).total; }
size_t dgr_binning_bytes(uint64_t cap, int32_t H, int32_t W) { (void)H; (void)W; return BinningLayout(cap).total; }

int dgr_forward_preprocess(const DgrSettings *s, const DgrGaussians *g, void *geom_v, void *image_v, int32_t *radii, void *stream) {
    NvtxRange nvtx_("dgr_forward_preprocess");
    if (int e = check_settings(s)) return e;
    if (int e = check_gaussians(s, g)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v, *image = (char *)image_v;
    if (!geom || !image) return fail(-1, "geom / image scratch is NULL");
    GeomLayout L(g->P, s->image_height, s->image_width);
    ImageLayout IL(s->image_height, s->image_width);
    if ((size_t)L.tiles * 4 > kMaxTileSmem) return fail(-3, "image has more than 51200 tiles (16x16): not supported");
    // one memset: the work block + the per-tile instance totals the preprocess kernel adds into + the backward cost accumulators
    DGR_CUDA(cudaMemsetAsync(image + IL.off_work, 0, (IL.off_cost - IL.off_work) + (size_t)L.tiles * 8 * 4, st));
    if (g->P > 0) {
        if (!radii) return fail(-1, "radii is NULL");
        DGR_KERNEL("preprocess_fwd", st, s->debug, DGR_DISPATCH(launch_pre_fwd, s, g, radii, geom, L, reinterpret_cast<unsigned *>(image + IL.off_count), st));
    }
    return 0;
}

int dgr_forward_render(const DgrSettings *s, const DgrGaussians *g, void *geom_v, void *binning_v, uint64_t capacity,
                       void *image_v, const DgrImages *out, int32_t flags, uint64_t *counts_host, uint64_t ticket,
                       void *count_ready_event, void *stream) {
    NvtxRange nvtx_("dgr_forward_render");
    if (int e = check_settings(s)) return e;
    if (!g || !geom_v || !image_v || !out) return fail(-1, "NULL argument");
    if (!out->color || !out->depth || !out->alpha) return fail(-1, "output images must not be NULL");
    if (capacity > 0 && !binning_v) return fail(-1, "binning scratch is NULL");
    if (capacity >= 0xffffffffull) return fail(-3, "capacity above 2^32-1 tile instances not supported");
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v, *binning = (char *)binning_v, *image = (char *)image_v;
    const int H = s->image_height, W = s->image_width;
    GeomLayout GL(g->P, H, W);
    ImageLayout IL(H, W);
    BinningLayout BL(capacity);
    const int tiles = IL.gx * IL.gy;
    if ((size_t)tiles * 4 > kMaxTileSmem) return fail(-3, "image has more than 51200 tiles (16x16): not supported");
    uint2 *ranges = reinterpret_cast<uint2 *>(image + IL.off_ranges);
    unsigned *tile_count = reinterpret_cast<unsigned *>(image + IL.off_count);
    const unsigned *run_matrix = reinterpret_cast<const unsigned *>(geom + GL.off_runs);
    unsigned *n_contrib = reinterpret_cast<unsigned *>(image + IL.off_ncontrib);
    float *final_T = reinterpret_cast<float *>(image + IL.off_finalT);
    GeomHeader *hdr = reinterpret_cast<GeomHeader *>(geom);
    const Rec *rec = reinterpret_cast<const Rec *>(geom + GL.off_rec);
    unsigned long long *keys = binning ? reinterpret_cast<unsigned long long *>(binning + BL.off_keys) : nullptr;
    unsigned *ids = binning ? reinterpret_cast<unsigned *>(binning + BL.off_ids) : nullptr;
    Rec *recs = binning ? reinterpret_cast<Rec *>(binning + BL.off_rec) : nullptr;

    unsigned *tile_order = reinterpret_cast<unsigned *>(image + IL.off_order);
    unsigned *big_list = reinterpret_cast<unsigned *>(image + IL.off_biglist);
    TileWork *work = reinterpret_cast<TileWork *>(image + IL.off_work);
    DevInfo *dv = dev_info();
    if (!dv) return -1;
    (void)(flags & DGR_FLAG_RERUN);           // a re-run only repeats the scan from the (still valid) per-tile totals
    // (a re-run with corrected guesses repeats everything from here: per-tile totals and run matrix are still valid; what the
    //  first, clipped render pass filed as backward costs is not)
    if (flags & DGR_FLAG_RERUN) {
        DGR_CUDA(cudaMemsetAsync(image + IL.off_cost, 0, (size_t)tiles * 8 * 4, st));
        DGR_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(work) + offsetof(TileWork, cost_bpt), 0, sizeof(TileWork) - offsetof(TileWork, cost_bpt), st));
    }
    {
        // instance emission; its extra block publishes ranges / order / counts (capacity 0: only that block has work to do)
        const size_t smem = (size_t)tiles * 4;
        const int nb = (g->P > 0 && capacity > 0) ? GL.nblocks : 0;
        DGR_KERNEL("emit_instances", st, s->debug,
                   launch_k(emit_instances_kernel, dim3(nb + 1), dim3(kPreThreads), smem, st, !(flags & DGR_FLAG_RERUN), g->P, IL.gx, tiles, GL.iters, nb, rec,
                            reinterpret_cast<const unsigned *>(geom + GL.off_touched), (const unsigned *)tile_count, (unsigned long long)capacity,
                            run_matrix, keys, ranges, hdr, tile_order, reinterpret_cast<uint2 *>(image + IL.off_oranges), work, big_list,
                            (volatile unsigned long long *)(ticket ? counts_host : nullptr), (unsigned long long)ticket));
    }
    if (!ticket) {            // copy + event between the kernels (this also ends the chain of programmatic dependent launches here)
        if (counts_host)      // { n_instances, n_big_tiles }
            DGR_CUDA(cudaMemcpyAsync(counts_host, geom, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (count_ready_event) DGR_CUDA(cudaEventRecord((cudaEvent_t)count_ready_event, st));
    }
    // Record staging of the render kernels.  A sorted copy of the records (written by the sort kernel, read with one contiguous
    // bulk-TMA copy per 32 records) is the fast path while that copy lives in L2; once it is several times the L2 size the
    // sort kernel's gather + write becomes the long pole (2M / 1600^2: 457 vs 164 us) and staging by id wins although a
    // chunk then takes 32 small bulk copies (B200: 78 vs 21 ns per chunk per SM, profiles/r2_tma_gather_probe.txt).
    const int lz = g_lazy.load();
    const bool lazy = lz == 2 || (lz == 0 && (size_t)capacity * sizeof(Rec) > 2 * dv->l2_bytes);
    if (g->P > 0 && capacity > 0) {
        if (g_sort_fine.load()) {
            const int sort_grid = min(tiles, persistent_grid(dv, tile_sort_gather_kernel<kSortSmallBucketsFine>, kSortSmallThreads, SmSF::bytes));
            DGR_KERNEL("tile_sort_gather", st, s->debug,
                       launch_k(tile_sort_gather_kernel<kSortSmallBucketsFine>, dim3(sort_grid), dim3(kSortSmallThreads), SmSF::bytes, st, true, (const TileWork *)work,
                                reinterpret_cast<const uint2 *>(image + IL.off_oranges), keys, rec, ids, recs, lazy ? 0 : 1));
        } else {
            const int sort_grid = min(tiles, persistent_grid(dv, tile_sort_gather_kernel<kSortSmallBuckets>, kSortSmallThreads, SmS::bytes));
            DGR_KERNEL("tile_sort_gather", st, s->debug,
                       launch_k(tile_sort_gather_kernel<kSortSmallBuckets>, dim3(sort_grid), dim3(kSortSmallThreads), SmS::bytes, st, true, (const TileWork *)work,
                                reinterpret_cast<const uint2 *>(image + IL.off_oranges), keys, rec, ids, recs, lazy ? 0 : 1));
        }
        if (flags & DGR_FLAG_BIG_TILES)
            DGR_KERNEL("tile_sort_gather_big", st, s->debug,
                       launch_k(tile_sort_gather_big_kernel, dim3(dv->big_grid), dim3(kSortBigThreads), SmB::bytes, st, true, (const TileWork *)work,
                                (const unsigned *)big_list, (const uint2 *)ranges, keys, rec, ids, recs, lazy ? 0 : 1));
    }
    // persistent forward render: every (tile, sub-tile) is a work item, handed out heaviest tile first
    const int ppl = g_ppl_fwd.load(), uf = g_u_fwd.load();
    CostOrder co{};
    if (g_cost_order.load()) {
        co.cost_acc = reinterpret_cast<unsigned *>(image + IL.off_cost); co.cls_count = work->cls_count;
        co.cls_items = reinterpret_cast<unsigned *>(image + IL.off_clsitems); co.cost_bpt = &work->cost_bpt;
        co.tiles = tiles; co.bwd_ppl = g_ppl_bwd.load();
    }
#define DGR_RENDER_FWD_L(PPL_, U_, L_)                                                                          \
    do {                                                                                                        \
        const int items_ = tiles * SubTile<PPL_>::kPerTile;                                                     \
        int grid_ = min(persistent_grid(dv, render_fwd_kernel<PPL_, U_, L_>, kRenderThreads, 0), (items_ + kRenderWarps - 1) / kRenderWarps); \
        if (g_cta_fwd.load() > 0) grid_ = min(grid_, g_cta_fwd.load() * dv->sms);                               \
        DGR_KERNEL("render_fwd", st, s->debug,                                                                  \
                   launch_k(render_fwd_kernel<PPL_, U_, L_>, dim3((unsigned)grid_), dim3(kRenderThreads), 0, st, true, H, W, IL.gx,   \
                            (const unsigned *)tile_order, reinterpret_cast<const uint2 *>(image + IL.off_oranges),         \
                            (const unsigned *)&work->n_nonempty, (unsigned)items_, &work->fwd_next, g_two_ended.load(), dv->sms, co, \
                            &work->lazy, (const Rec *)recs, rec, (const unsigned *)ids,                                      \
                            s->bg, out->color, out->depth, out->alpha, n_contrib, final_T));                     \
    } while (0)
#define DGR_RENDER_FWD(PPL_, U_) do { if (lazy) DGR_RENDER_FWD_L(PPL_, U_, true); else DGR_RENDER_FWD_L(PPL_, U_, false); } while (0)
    if (ppl == 4) DGR_RENDER_FWD(4, 1);
    else if (ppl == 2) DGR_RENDER_FWD(2, 2);
    else { if (uf == 4) DGR_RENDER_FWD(1, 4); else DGR_RENDER_FWD(1, 1); }      // default <1,1>: 58 registers, 8 CTAs per SM, no spill
#undef DGR_RENDER_FWD_L
#undef DGR_RENDER_FWD
    return 0;
}

static int check_push(const DgrPeerPush *p, int64_t P) {
    if (!p) return 0;
    if (p->world < 1 || p->world > kMaxPeers || p->rank < 0 || p->rank >= p->world) return fail(-1, "push: bad world / rank");
    if (p->gaussians_per_owner <= 0 || p->gaussians_per_owner % kPreThreads != 0) return fail(-1, "push: gaussians_per_owner must be a positive multiple of 256");
    if (p->gaussians_per_owner * p->world < P) return fail(-1, "push: gaussians_per_owner * world < P");
    if (p->delta_floats[p->rank] != 0) return fail(-1, "push: delta_floats[rank] must be 0");
    return 0;
}

static int flat_segs(int32_t n_seg, const int64_t *seg_off, const int32_t *seg_stride, FlatSegs *out) {
    if (n_seg < 1 || n_seg > 8 || !seg_off || !seg_stride) return fail(-1, "1..8 flat segments expected");
    out->n = n_seg;
    for (int k = 0; k < n_seg; k++) {
        if (seg_off[k] < 0 || seg_stride[k] < 1) return fail(-1, "bad flat segment");
        out->off[k] = seg_off[k]; out->stride[k] = seg_stride[k];
    }
    return 0;
}

int dgr_backward(const DgrSettings *s, const DgrGaussians *g, void *geom_v, const void *binning_v, uint64_t capacity,
                 const void *image_v, const int32_t *radii, const float *out_alpha, const DgrImageGrads *gin,
                 const DgrGaussianGrads *gout, void *stream) {
    NvtxRange nvtx_("dgr_backward");
    (void)out_alpha;
    if (int e = check_settings(s)) return e;
    if (int e = check_gaussians(s, g)) return e;
    if (!geom_v || !image_v || !gin || !gout) return fail(-1, "NULL argument");
    if (g->P == 0) return 0;
    if (!radii) return fail(-1, "radii is NULL");
    if (int e = check_push(gout->push, g->P)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    char *geom = (char *)geom_v;
    const char *binning = (const char *)binning_v, *image = (const char *)image_v;
    const int H = s->image_height, W = s->image_width;
    GeomLayout GL(g->P, H, W);
    ImageLayout IL(H, W);
    BinningLayout BL(capacity);
    const int tiles = IL.gx * IL.gy;
    float *grad_rec = reinterpret_cast<float *>(geom + GL.off_gradrec);
    // one memset: the work counter of the persistent backward render kernel + the per-Gaussian moment accumulators
    DGR_CUDA(cudaMemsetAsync(geom + GL.off_bwdwork, 0, (GL.off_gradrec - GL.off_bwdwork) + (size_t)g->P * kGradRecFloats * 4, st));
    if (binning && capacity > 0) {
        DevInfo *dv = dev_info();
        if (!dv) return -1;
        const TileWork *work = reinterpret_cast<const TileWork *>(image + IL.off_work);
        unsigned long long *bwd_next = reinterpret_cast<unsigned long long *>(geom + GL.off_bwdwork);
        const int ppl = g_ppl_bwd.load(), ub = g_u_bwd.load();
        CostOrder co{};
        if (g_cost_order.load()) {            // (read-only here; the kernel checks that the forward grouped its costs for this shape)
            TileWork *wk = const_cast<TileWork *>(work);
            co.cost_acc = const_cast<unsigned *>(reinterpret_cast<const unsigned *>(image + IL.off_cost)); co.cls_count = wk->cls_count;
            co.cls_items = const_cast<unsigned *>(reinterpret_cast<const unsigned *>(image + IL.off_clsitems)); co.cost_bpt = &wk->cost_bpt;
            co.tiles = tiles; co.bwd_ppl = ppl == 1 ? 1 : 2;
        }
#define DGR_RENDER_BWD_L(PPL_, U_, L_)                                                                                        \
    do {                                                                                                                   \
        const size_t smem_ = sizeof(BwdSmem<PPL_>) * kRenderWarps;                                                         \
        int grid_ = min(persistent_grid(dv, render_bwd_kernel<PPL_, U_, L_>, kRenderThreads, smem_),                       \
                        (tiles * SubTile<PPL_>::kPerTile + kRenderWarps - 1) / kRenderWarps);                              \
        if (g_cta_bwd.load() > 0) grid_ = min(grid_, g_cta_bwd.load() * dv->sms);                                          \
        DGR_KERNEL("render_bwd", st, s->debug,                                                                             \
                   launch_k(render_bwd_kernel<PPL_, U_, L_>, dim3((unsigned)grid_), dim3(kRenderThreads), smem_, st, false, H, W, IL.gx, \
                            reinterpret_cast<const unsigned *>(image + IL.off_order), &work->n_nonempty, bwd_next, g_two_ended.load(), dv->sms, co, \
                            reinterpret_cast<const uint2 *>(image + IL.off_oranges),                                       \
                            reinterpret_cast<const Rec *>(binning + BL.off_rec), reinterpret_cast<const Rec *>(geom + GL.off_rec), \
                            (const unsigned *)&work->lazy, reinterpret_cast<const unsigned *>(binning + BL.off_ids),          \
                            s->bg, reinterpret_cast<const float *>(image + IL.off_finalT),                                 \
                            reinterpret_cast<const unsigned *>(image + IL.off_ncontrib), gin->dL_dcolor, gin->dL_ddepth, gin->dL_dalpha, grad_rec)); \
    } while (0)
#define DGR_RENDER_BWD(PPL_, U_) do { if (lazy) DGR_RENDER_BWD_L(PPL_, U_, true); else DGR_RENDER_BWD_L(PPL_, U_, false); } while (0)
        // the record staging follows the rule the forward applied to the same capacity (dgr_forward_render)
        const int lz = g_lazy.load();
        const bool lazy = lz == 2 || (lz == 0 && (size_t)capacity * sizeof(Rec) > 2 * dv->l2_bytes);
        if (ppl == 1 && ub != 1 && !lazy && !g_rowsum.load()) {       // A/B: the default shape with the per-pixel step 2
#define DGR_BWD_OLD_ render_bwd_kernel<1, 2, false, false>
            const size_t smem_ = sizeof(BwdSmem<1>) * kRenderWarps;
            int grid_ = min(persistent_grid(dv, DGR_BWD_OLD_, kRenderThreads, smem_), (tiles * SubTile<1>::kPerTile + kRenderWarps - 1) / kRenderWarps);
            if (g_cta_bwd.load() > 0) grid_ = min(grid_, g_cta_bwd.load() * dv->sms);
            DGR_KERNEL("render_bwd", st, s->debug,
                       launch_k(DGR_BWD_OLD_, dim3((unsigned)grid_), dim3(kRenderThreads), smem_, st, false, H, W, IL.gx,
                                reinterpret_cast<const unsigned *>(image + IL.off_order), &work->n_nonempty, bwd_next, g_two_ended.load(), dv->sms, co,
                                reinterpret_cast<const uint2 *>(image + IL.off_oranges),
                                reinterpret_cast<const Rec *>(binning + BL.off_rec), reinterpret_cast<const Rec *>(geom + GL.off_rec),
                                (const unsigned *)&work->lazy, reinterpret_cast<const unsigned *>(binning + BL.off_ids),
                                s->bg, reinterpret_cast<const float *>(image + IL.off_finalT),
                                reinterpret_cast<const unsigned *>(image + IL.off_ncontrib), gin->dL_dcolor, gin->dL_ddepth, gin->dL_dalpha, grad_rec));
#undef DGR_BWD_OLD_
        } else
        if (ppl == 1) { if (ub == 1) DGR_RENDER_BWD(1, 1); else DGR_RENDER_BWD(1, 2); }     // default <1,2> (in-pipeline sweep, profiles/r2_sweep.txt)
        else { if (ub == 2) DGR_RENDER_BWD(2, 2); else DGR_RENDER_BWD(2, 1); }
#undef DGR_RENDER_BWD_L
#undef DGR_RENDER_BWD
    }
    DGR_KERNEL("preprocess_bwd", st, s->debug, DGR_DISPATCH(launch_pre_bwd, s, g, radii, reinterpret_cast<const unsigned *>(geom + GL.off_touched), grad_rec, gout, st));
    return 0;
}

#include "dgr_collective_api.cuh"      // dgr_peer_*: the multi-GPU entry points (not part of the single-GPU step)

size_t dgr_knn_scratch_bytes(int32_t P) { return KnnLayout(P).total; }

int dgr_dist_cuda2(int32_t P, const float *points, float *mean_dists, void *scratch_v, void *stream) {
    NvtxRange nvtx_("dgr_dist_cuda2");
    if (P < 0) return fail(-1, "P < 0");
    if (P == 0) return 0;
    if (!points || !mean_dists || !scratch_v) return fail(-1, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    char *scratch = (char *)scratch_v;
    KnnLayout L(P);
    KnnGrid *grid = reinterpret_cast<KnnGrid *>(scratch + L.off_grid);
    unsigned *cell_start = reinterpret_cast<unsigned *>(scratch + L.off_start);
    unsigned *cell_fill = reinterpret_cast<unsigned *>(scratch + L.off_fill);
    unsigned *sums = reinterpret_cast<unsigned *>(scratch + L.off_sums);
    unsigned *cid = reinterpret_cast<unsigned *>(scratch + L.off_cid);
    float4 *sorted = reinterpret_cast<float4 *>(scratch + L.off_sorted);
    DGR_CUDA(cudaMemsetAsync(grid, 0, sizeof(KnnGrid), st));
    DGR_CUDA(cudaMemsetAsync(grid, 0xff, 3 * sizeof(unsigned), st));                      // lo[] = +inf in the ordered encoding
    DGR_CUDA(cudaMemsetAsync(cell_start, 0, (L.cap + 1) * 4, st));
    DGR_CUDA(cudaMemsetAsync(cell_fill, 0, L.cap * 4, st));
    const int nb = (P + 255) / 256;
    const int nb_bbox = nb < 4 * (dev_info() ? dev_info()->sms : 148) ? nb : 4 * (dev_info() ? dev_info()->sms : 148);
    const size_t n_scan = L.cap + 1;
    DGR_KERNEL("knn_bbox", st, 0, knn_bbox_kernel<<<nb_bbox, 256, 0, st>>>(P, points, grid));
    DGR_KERNEL("knn_grid", st, 0, knn_grid_kernel<<<1, 1, 0, st>>>(P, (unsigned)L.cap, grid));
    DGR_KERNEL("knn_count", st, 0, knn_count_kernel<<<nb, 256, 0, st>>>(P, points, grid, cell_start, cid));
    DGR_KERNEL("knn_scan_sums", st, 0, knn_scan_sums_kernel<<<(unsigned)L.nblk, 1024, 0, st>>>(cell_start, n_scan, sums));
    DGR_KERNEL("knn_scan_top", st, 0, knn_scan_top_kernel<<<1, 1024, 0, st>>>(sums, L.nblk));
    DGR_KERNEL("knn_scan_apply", st, 0, knn_scan_apply_kernel<<<(unsigned)L.nblk, 1024, 0, st>>>(cell_start, n_scan, sums));
    DGR_KERNEL("knn_scatter", st, 0, knn_scatter_kernel<<<nb, 256, 0, st>>>(P, points, cid, cell_start, cell_fill, sorted));
    DGR_KERNEL("knn_query", st, 0, knn_query_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, grid, cell_start, sorted, mean_dists));
    return 0;
}

size_t dgr_fields_scratch_bytes(int32_t P, int32_t num_blocks) { return FieldsLayout(P, num_blocks > 0 ? num_blocks : 1).total; }

int dgr_extract_fields(int32_t P, const float *xyz, const float *opacity_raw, const float *scaling_raw, const float *rotation_raw,
                       int32_t resolution, int32_t num_blocks, float relax_ratio, float *occ, float *center_scale, void *scratch_v,
                       void *stream) {
    NvtxRange nvtx_("dgr_extract_fields");
    if (P < 0 || resolution < 2 || num_blocks < 1) return fail(-1, "bad P / resolution / num_blocks");
    if (resolution % num_blocks != 0) return fail(-2, "resolution must be a multiple of num_blocks");
    if (num_blocks > 64) return fail(-3, "num_blocks above 64 not supported");
    const int split = resolution / num_blocks;
    const int V = split * split * split;
    if (V > 16 * kFieldThreads) return fail(-3, "more than 4096 voxels per block not supported");
    if (!occ || !scratch_v || (P > 0 && (!xyz || !opacity_raw || !scaling_raw || !rotation_raw))) return fail(-1, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    char *scratch = (char *)scratch_v;
    FieldsLayout L(P, num_blocks);
    FieldsHeader *hdr = reinterpret_cast<FieldsHeader *>(scratch + L.off_hdr);
    unsigned *cell_start = reinterpret_cast<unsigned *>(scratch + L.off_start);
    unsigned *cell_fill = reinterpret_cast<unsigned *>(scratch + L.off_fill);
    unsigned *cid = reinterpret_cast<unsigned *>(scratch + L.off_cid);
    FieldRec *rec = reinterpret_cast<FieldRec *>(scratch + L.off_rec), *sorted = reinterpret_cast<FieldRec *>(scratch + L.off_sorted);
    DGR_CUDA(cudaMemsetAsync(hdr, 0, sizeof(FieldsHeader), st));
    DGR_CUDA(cudaMemsetAsync(hdr, 0xff, 3 * sizeof(unsigned), st));
    DGR_CUDA(cudaMemsetAsync(cell_start, 0, (L.cells + 1) * 4, st));
    DGR_CUDA(cudaMemsetAsync(cell_fill, 0, L.cells * 4, st));
    const int nbk = P > 0 ? (P + 255) / 256 : 1;
    if (P > 0) {
        const int nb_bbox = nbk < 4 * (dev_info() ? dev_info()->sms : 148) ? nbk : 4 * (dev_info() ? dev_info()->sms : 148);
        DGR_KERNEL("fields_bbox", st, 0, fields_bbox_kernel<<<nb_bbox, 256, 0, st>>>(P, xyz, opacity_raw, hdr));
    }
    DGR_KERNEL("fields_prep", st, 0, fields_prep_kernel<<<nbk, 256, 0, st>>>(P, xyz, opacity_raw, scaling_raw, rotation_raw, hdr, num_blocks,
                                                                              rec, cid, cell_start, center_scale));
    DGR_KERNEL("fields_scan", st, 0, fields_scan_kernel<<<1, 1024, 0, st>>>(cell_start, (int)(L.cells + 1)));
    if (P > 0)
        DGR_KERNEL("fields_scatter", st, 0, fields_scatter_kernel<<<nbk, 256, 0, st>>>(P, cid, cell_start, cell_fill, rec, sorted));
    // reference: vmin -= block_size * relax_ratio with python floats (gs_renderer.py:222,264-265) -> one float32 operand
    const float grow = (float)((2.0 / (double)num_blocks) * (double)relax_ratio);
    const unsigned grid = (unsigned)L.cells;
    const int vpt = (V + kFieldThreads - 1) / kFieldThreads;
#define DGR_FIELDS(VPT_) DGR_KERNEL("fields_eval", st, 0, fields_eval_kernel<VPT_><<<grid, kFieldThreads, 0, st>>>(resolution, num_blocks, split, grow, cell_start, sorted, occ))
    if (vpt <= 1) DGR_FIELDS(1); else if (vpt <= 2) DGR_FIELDS(2); else if (vpt <= 4) DGR_FIELDS(4); else if (vpt <= 8) DGR_FIELDS(8); else DGR_FIELDS(16);
#undef DGR_FIELDS
    return 0;
}

int dgr_adam_step(const DgrAdamGroup *groups, int32_t n_groups, double beta1, double beta2, double eps, void *stream) {
    NvtxRange nvtx_("dgr_adam_step");
    if (!groups || n_groups < 1 || n_groups > kAdamMaxGroups) return fail(-1, "dgr_adam_step: 1..8 groups");
    AdamGroups G;
    G.n_groups = n_groups;
    unsigned long long run = 0;
    for (int k = 0; k < n_groups; k++) {
        const DgrAdamGroup &g = groups[k];
        if (g.n > 0 && (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq)) return fail(-1, "dgr_adam_step: NULL tensor in a non-empty group");
        G.param[k] = g.param; G.grad[k] = g.grad; G.m[k] = g.exp_avg; G.v[k] = g.exp_avg_sq; G.n[k] = g.n;
        if (g.step < 1) return fail(-1, "dgr_adam_step: a tensor's step counts from 1");
        const double bc1 = 1.0 - pow(beta1, (double)g.step), bc2 = 1.0 - pow(beta2, (double)g.step);
        G.step_size[k] = (float)((double)g.lr / bc1);
        G.inv_sqrt_bc2[k] = (float)(1.0 / sqrt(bc2));
        G.start[k] = run; run += (g.n + 3) / 4;
    }
    for (int k = n_groups; k <= kAdamMaxGroups; k++) G.start[k] = run;
    G.start[n_groups] = run;
    if (run == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long blocks = (run + 255) / 256;
    const unsigned long long cap = (unsigned long long)(dev_info() ? dev_info()->sms : 148) * 16;
    if (blocks > cap) blocks = cap;
    DGR_KERNEL("adam_multi", st, 0, adam_multi_kernel<<<(unsigned)blocks, 256, 0, st>>>(G, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps));
    return 0;
}

size_t dgr_densify_scratch_bytes(int32_t P) { return DensifyLayout(P).total; }

int dgr_densify_plan(int32_t P, const float *xyz_gradient_accum, const float *denom, const float *opacity_raw, const float *scaling_raw,
                     float grad_threshold, float dense_extent, float min_opacity, float max_world, int32_t use_world,
                     void *scratch_v, uint32_t *counts_host, void *stream) {
    NvtxRange nvtx_("dgr_densify_plan");
    if (P < 0) return fail(-1, "P < 0");
    if (!scratch_v || (P > 0 && (!xyz_gradient_accum || !denom || !opacity_raw || !scaling_raw))) return fail(-1, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    char *scratch = (char *)scratch_v;
    DensifyLayout L(P);
    unsigned *counts = reinterpret_cast<unsigned *>(scratch + L.off_counts), *offsets = reinterpret_cast<unsigned *>(scratch + L.off_offsets);
    unsigned *totals = reinterpret_cast<unsigned *>(scratch + L.off_totals);
    DensifyParams prm{grad_threshold, dense_extent, min_opacity, max_world, use_world};
    if (P > 0)
        DGR_KERNEL("densify_classify", st, 0,
                   densify_classify_kernel<<<L.nblk, kDensifyThreads, 0, st>>>(P, xyz_gradient_accum, denom, opacity_raw, scaling_raw, prm,
                                                                                reinterpret_cast<unsigned char *>(scratch + L.off_class), counts));
    else
        DGR_CUDA(cudaMemsetAsync(counts, 0, 16, st));
    DGR_KERNEL("densify_offsets", st, 0, densify_offsets_kernel<<<1, 1024, 0, st>>>(P > 0 ? L.nblk : 1, counts, offsets, totals));
    if (counts_host) DGR_CUDA(cudaMemcpyAsync(counts_host, totals, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    return 0;
}

int dgr_densify_apply(int32_t P, const DgrDensifyTensors *t, const float *noise, const void *scratch_v, void *stream) {
    NvtxRange nvtx_("dgr_densify_apply");
    if (P < 0 || !t || !scratch_v) return fail(-1, "bad argument");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const char *scratch = (const char *)scratch_v;
    DensifyLayout L(P);
    DensifyTensors T;
    for (int k = 0; k < kDensifyTensors; k++) {
        T.in[k] = t->in[k]; T.m_in[k] = t->exp_avg_in[k]; T.v_in[k] = t->exp_avg_sq_in[k];
        T.out[k] = t->out[k]; T.m_out[k] = t->exp_avg_out[k]; T.v_out[k] = t->exp_avg_sq_out[k];
        T.width[k] = t->width[k];
        if (t->width[k] < 0) return fail(-1, "negative row width");
        if (t->width[k] > 0 && (!T.in[k] || !T.m_in[k] || !T.v_in[k])) return fail(-1, "NULL input tensor");
    }
    if (T.width[0] != 3 || T.width[3] != 1 || T.width[4] != 3 || T.width[5] != 4) return fail(-1, "row widths must be xyz 3, opacity 1, scaling 3, rotation 4");
    DGR_KERNEL("densify_apply", st, 0,
               densify_apply_kernel<<<L.nblk, kDensifyThreads, 0, st>>>(P, reinterpret_cast<const unsigned char *>(scratch + L.off_class),
                                                                        reinterpret_cast<const unsigned *>(scratch + L.off_offsets),
                                                                        reinterpret_cast<const unsigned *>(scratch + L.off_totals), T, noise));
    return 0;
}

int dgr_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                     uint8_t *present, void *stream) {
    NvtxRange nvtx_("dgr_mark_visible");
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(-1, "bad argument");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DGR_KERNEL("mark_visible", st, 0, mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, viewmatrix, present));
    return 0;
}

int dgr_debug_geom(int32_t P, int32_t H, int32_t W, const void *geom_v, float *mean_px, float *depth, float *conic, float *rgb,
                   int32_t *aabb, uint32_t *tiles_touched, void *stream) {
    if (P <= 0) return 0;
    if (!geom_v) return fail(-1, "geom is NULL");
    const char *geom = (const char *)geom_v;
    GeomLayout L(P, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    DGR_KERNEL("debug_geom", st, 0,
               debug_geom_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, reinterpret_cast<const Rec *>(geom + L.off_rec),
                                                                reinterpret_cast<const unsigned *>(geom + L.off_touched), mean_px,
                                                                depth, conic, rgb, aabb, tiles_touched));
    return 0;
}

}  // extern "C"
