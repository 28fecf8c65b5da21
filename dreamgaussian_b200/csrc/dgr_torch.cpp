// dgr_torch.cpp — compiled host layer over the C ABI (include/dgr_b200.h) for PyTorch callers.
//
// The ctypes path (dreamgaussian_b200/rasterizer.py) costs ~0.25 ms of Python per forward+backward — as much as the GPU
// work at 100k Gaussians (tools/host_overhead.py).  This module does the same calls from C++: tensor checks, scratch
// allocation through the caching allocator, the speculative-capacity protocol of dgr_forward_render (count event, re-run
// on a low guess) and the backward, ~20 µs of host time each.  It adds NO arithmetic: every kernel still lives in
// libdgr_b200.so, which this module links.  Mirrors rasterize_gaussians / rasterize_gaussians_backward of the reference
// package's ext.cpp [EXT] (called from /root/reference/gs_renderer.py:800-809 through GaussianRasterizer).
#include <torch/extension.h>
#include <cstdint>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>

#include <mutex>
#include <unordered_map>

#include "../../include/dgr_b200.h"

namespace {

using torch::Tensor;
using OptTensor = c10::optional<Tensor>;

const float *fptr(const OptTensor &t) { return (t.has_value() && t->defined() && t->numel() > 0) ? t->data_ptr<float>() : nullptr; }
float *fptr_mut(const OptTensor &t) { return (t.has_value() && t->defined() && t->numel() > 0) ? t->data_ptr<float>() : nullptr; }

Tensor dev_f32(const Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda(), name, " is on ", t.device(), ": this rasterizer has no CPU path (it runs as sm_100a CUDA kernels in libdgr_b200.so)");
    Tensor r = t.scalar_type() == torch::kFloat32 ? t : t.to(torch::kFloat32);
    r = r.contiguous();
    // the kernels read rotations / SH rows with 128-bit loads and stage inputs with bulk TMA: a view that starts off a 16-byte
    // boundary gets its own (aligned) allocation
    if (r.numel() > 0 && (reinterpret_cast<uintptr_t>(r.data_ptr()) & 15u) != 0) r = r.clone(at::MemoryFormat::Contiguous);
    return r;
}
OptTensor opt_f32(const OptTensor &t, const char *name) {
    if (!t.has_value() || !t->defined() || (t->numel() == 0 && t->dim() <= 1)) return c10::nullopt;
    return dev_f32(*t, name);
}
void check(int rc) { TORCH_CHECK(rc == 0, "libdgr_b200: ", dgr_last_error()); }

struct State {                       // what one forward leaves behind for its backward (keeps every tensor alive)
    DgrSettings s;
    DgrGaussians g;
    Tensor bg, view, proj, campos;
    Tensor means3D, opac;
    OptTensor sh, colors, scales, rots, cov3D, sh_rest;
    Tensor geom, binning, image, radii;      // NOT the differentiable outputs: State -> output -> grad_fn -> ctx -> State would be an
                                             // uncollectable cycle through C++ reference counts (radii is non-differentiable: no grad_fn)
    int64_t capacity = 0, n_inst = 0;
    int device = 0;
};

struct Hint { int64_t cap; bool big; };
std::mutex g_mu;
std::unordered_map<uint64_t, Hint> g_hints;
uint64_t hint_key(int dev, int64_t P, int H, int W) { return ((uint64_t)dev << 58) ^ ((uint64_t)P << 28) ^ ((uint64_t)H << 14) ^ (uint64_t)W; }

struct SyncObjs { Tensor counts; void *event = nullptr; uint64_t ticket = 0; };
SyncObjs &sync_objs(int dev) {
    thread_local std::unordered_map<int, SyncObjs> per_dev;
    SyncObjs &o = per_dev[dev];
    if (!o.event) {
        o.counts = torch::zeros({4}, torch::dtype(torch::kInt64)).pin_memory();
        o.event = dgr_event_create();
        TORCH_CHECK(o.event, "libdgr_b200: could not create a CUDA event");
    }
    return o;
}

std::tuple<Tensor, Tensor, Tensor, Tensor, std::shared_ptr<State>>
forward(int64_t H, int64_t W, double tanfovx, double tanfovy, double scale_modifier, int64_t sh_degree, bool prefiltered, bool debug,
        const Tensor &bg, const Tensor &view, const Tensor &proj, const Tensor &campos, const Tensor &means3D, const OptTensor &sh,
        const OptTensor &colors, const Tensor &opac, const OptTensor &scales, const OptTensor &rots, const OptTensor &cov3D,
        const OptTensor &sh_rest, bool activations) {
    auto st = std::make_shared<State>();
    st->means3D = dev_f32(means3D, "means3D");
    const c10::cuda::CUDAGuard guard(st->means3D.device());
    st->device = st->means3D.get_device();
    st->opac = dev_f32(opac, "opacities");
    st->bg = dev_f32(bg, "bg"); st->view = dev_f32(view, "viewmatrix"); st->proj = dev_f32(proj, "projmatrix"); st->campos = dev_f32(campos, "campos");
    TORCH_CHECK_VALUE(st->bg.numel() == 3 && st->view.numel() == 16 && st->proj.numel() == 16 && st->campos.numel() == 3,
                "bg/campos must have 3 elements, viewmatrix/projmatrix 16");
    st->sh = opt_f32(sh, "shs"); st->colors = opt_f32(colors, "colors_precomp"); st->scales = opt_f32(scales, "scales");
    st->rots = opt_f32(rots, "rotations"); st->cov3D = opt_f32(cov3D, "cov3D_precomp");
    st->sh_rest = activations ? opt_f32(sh_rest, "features_rest") : c10::nullopt;
    if (st->sh_rest.has_value() && st->sh_rest->numel() == 0) st->sh_rest = c10::nullopt;
    TORCH_CHECK_VALUE(st->means3D.dim() == 2 && st->means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    const int64_t P = st->means3D.size(0);
    int64_t M = 0;
    if (st->sh.has_value()) {
        TORCH_CHECK_VALUE(st->sh->dim() == 3 && st->sh->size(0) == P && st->sh->size(2) == 3, "shs must have dimensions (num_points, num_coeffs, 3)");
        M = st->sh->size(1);
    }
    if (activations) {
        TORCH_CHECK_VALUE(st->sh.has_value() && st->sh->size(1) == 1, "features_dc must have dimensions (num_points, 1, 3)");
        if (st->sh_rest.has_value()) {
            TORCH_CHECK_VALUE(st->sh_rest->dim() == 3 && st->sh_rest->size(0) == P && st->sh_rest->size(2) == 3, "features_rest must have dimensions (num_points, num_coeffs - 1, 3)");
            M = 1 + st->sh_rest->size(1);
        }
    }
    TORCH_CHECK_VALUE(!st->colors.has_value() || st->colors->numel() == P * 3, "colors_precomp must have dimensions (num_points, 3)");
    TORCH_CHECK_VALUE(!st->scales.has_value() || st->scales->numel() == P * 3, "scales must have dimensions (num_points, 3)");
    TORCH_CHECK_VALUE(!st->rots.has_value() || st->rots->numel() == P * 4, "rotations must have dimensions (num_points, 4)");
    TORCH_CHECK_VALUE(!st->cov3D.has_value() || st->cov3D->numel() == P * 6, "cov3D_precomp must have dimensions (num_points, 6)");
    TORCH_CHECK_VALUE(st->opac.numel() == P, "opacities must have dimensions (num_points, 1)");

    st->s = DgrSettings{(int32_t)H, (int32_t)W, (float)tanfovx, (float)tanfovy, (float)scale_modifier, (int32_t)sh_degree,
                        prefiltered ? 1 : 0, debug ? 1 : 0, st->bg.data_ptr<float>(), st->view.data_ptr<float>(), st->proj.data_ptr<float>(),
                        st->campos.data_ptr<float>()};
    st->g = DgrGaussians{(int32_t)P, (int32_t)M, st->means3D.numel() ? st->means3D.data_ptr<float>() : nullptr, fptr(st->sh), fptr(st->colors),
                         st->opac.numel() ? st->opac.data_ptr<float>() : nullptr, fptr(st->scales), fptr(st->rots), fptr(st->cov3D),
                         fptr(st->sh_rest), activations ? 1 : 0};
    const auto f32 = st->means3D.options();
    const auto u8 = f32.dtype(torch::kUInt8);
    Tensor img = torch::empty({5, H, W}, f32);               // color (3) | depth (1) | alpha (1): one allocation
    Tensor color = img.narrow(0, 0, 3), depth = img.narrow(0, 3, 1), alpha = img.narrow(0, 4, 1);
    st->radii = torch::empty({P}, f32.dtype(torch::kInt32));
    st->geom = torch::empty({(int64_t)dgr_geom_bytes((int32_t)P, (int32_t)H, (int32_t)W)}, u8);
    st->image = torch::empty({(int64_t)dgr_image_bytes((int32_t)H, (int32_t)W)}, u8);
    void *stream = c10::cuda::getCurrentCUDAStream(st->device).stream();
    check(dgr_forward_preprocess(&st->s, &st->g, st->geom.data_ptr(), st->image.data_ptr(), st->radii.data_ptr<int32_t>(), stream));
    DgrImages out{color.data_ptr<float>(), depth.data_ptr<float>(), alpha.data_ptr<float>(), st->radii.data_ptr<int32_t>()};
    Hint hint{std::max<int64_t>(65536, 16 * P), true};
    const uint64_t key = hint_key(st->device, P, (int)H, (int)W);
    { std::lock_guard<std::mutex> lk(g_mu); auto it = g_hints.find(key); if (it != g_hints.end()) hint = it->second; }
    SyncObjs &so = sync_objs(st->device);
    int64_t cap = hint.cap, n_inst = 0, n_big = 0;
    bool big = hint.big;
    int rerun = 0;
    while (true) {
        st->binning = torch::empty({(int64_t)dgr_binning_bytes((uint64_t)cap, (int32_t)H, (int32_t)W)}, u8);
        // ticket protocol: the scan kernel writes the counts and then the ticket into the pinned buffer; nothing but kernels
        // goes into the stream (programmatic dependent launches chain the whole forward) and the host polls the ticket
        const uint64_t ticket = ++so.ticket;
        volatile uint64_t *cnt = reinterpret_cast<volatile uint64_t *>(so.counts.data_ptr<int64_t>());
        check(dgr_forward_render(&st->s, &st->g, st->geom.data_ptr(), st->binning.data_ptr(), (uint64_t)cap, st->image.data_ptr(), &out,
                                 (big ? DGR_FLAG_BIG_TILES : 0) | rerun, reinterpret_cast<uint64_t *>(so.counts.data_ptr<int64_t>()), ticket, nullptr, stream));
        {
            pybind11::gil_scoped_release nogil;              // the wait for the early instance count
            uint64_t spins = 0;
            while (cnt[2] != ticket) {
                if ((++spins & 0xfffff) == 0) {              // every ~1M polls: has the stream failed or finished without the ticket?
                    const cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
                    if (q != cudaErrorNotReady && cnt[2] != ticket) {
                        TORCH_CHECK(q == cudaSuccess, "libdgr_b200: forward failed on the device: ", cudaGetErrorString(q));
                        TORCH_CHECK(false, "libdgr_b200: the instance count never arrived (is the count buffer device-mapped?)");
                    }
                }
            }
        }
        n_inst = so.counts.data_ptr<int64_t>()[0]; n_big = so.counts.data_ptr<int64_t>()[1];
        if (n_inst <= cap && (big || n_big == 0)) break;
        cap = std::max<int64_t>(cap, (int64_t)(n_inst * 1.25) + 4096);      // a guess was wrong: redo stage 2 (rare)
        big = big || n_big > 0;
        rerun = DGR_FLAG_RERUN;
    }
    { std::lock_guard<std::mutex> lk(g_mu); g_hints[key] = Hint{std::max<int64_t>((int64_t)(n_inst * 1.25) + 4096, 65536), n_big > 0}; }
    st->capacity = cap; st->n_inst = n_inst;
    return {color, st->radii, depth, alpha, st};
}

// grads: (means3D, means2D, shs | features_dc, colors, opacities, scales, rotations, cov3D, features_rest); undefined where not applicable
std::vector<Tensor> backward(const std::shared_ptr<State> &st, const OptTensor &gC, const OptTensor &gD, const OptTensor &gA, bool accumulate,
                             const std::vector<OptTensor> &out, const OptTensor &xyz_gradient_accum, const OptTensor &denom,
                             const OptTensor &max_radii2D, int64_t push_addr) {
    const c10::cuda::CUDAGuard guard(st->means3D.device());
    const int64_t P = st->g.P, M = st->g.M;
    const auto f32 = st->means3D.options();
    OptTensor c = gC.has_value() && gC->defined() ? OptTensor(dev_f32(*gC, "grad_color")) : c10::nullopt;
    OptTensor d = gD.has_value() && gD->defined() ? OptTensor(dev_f32(*gD, "grad_depth")) : c10::nullopt;
    OptTensor a = gA.has_value() && gA->defined() ? OptTensor(dev_f32(*gA, "grad_alpha")) : c10::nullopt;
    std::vector<Tensor> g(9);
    auto pick = [&](size_t i, bool wanted, c10::IntArrayRef shape) {
        if (i < out.size() && out[i].has_value() && out[i]->defined()) g[i] = *out[i];
        else if (wanted) g[i] = torch::empty(shape, f32);
    };
    const bool act = st->g.activations != 0;
    pick(0, true, {P, 3}); pick(1, true, {P, 3});
    pick(2, st->sh.has_value(), {P, act ? 1 : M, 3});
    pick(3, st->colors.has_value(), {P, 3});
    pick(4, true, {P, 1});
    pick(5, st->scales.has_value(), {P, 3}); pick(6, st->rots.has_value(), {P, 4}); pick(7, st->cov3D.has_value(), {P, 6});
    pick(8, act && st->sh_rest.has_value(), {P, M - 1, 3});
    auto gp = [&](size_t i) -> float * { return g[i].defined() && g[i].numel() > 0 ? g[i].data_ptr<float>() : nullptr; };
    DgrImageGrads gin{fptr(c), fptr(d), fptr(a)};
    DgrGaussianGrads gout{gp(0), gp(1), gp(2), gp(3), gp(4), gp(5), gp(6), gp(7), accumulate ? 1 : 0, gp(8),
                          fptr_mut(xyz_gradient_accum), fptr_mut(denom), fptr_mut(max_radii2D),
                          reinterpret_cast<const DgrPeerPush *>(push_addr)};      // host struct kept alive by the caller (0 = none)
    void *stream = c10::cuda::getCurrentCUDAStream(st->device).stream();
    check(dgr_backward(&st->s, &st->g, st->geom.data_ptr(), st->binning.data_ptr(), (uint64_t)st->capacity, st->image.data_ptr(),
                       st->radii.data_ptr<int32_t>(), nullptr, &gin, &gout, stream));
    return g;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    pybind11::class_<State, std::shared_ptr<State>>(m, "State")
        .def_readonly("capacity", &State::capacity)
        .def_readonly("num_rendered", &State::n_inst);
    m.def("forward", &forward, "both forward stages through the C ABI (no arithmetic here)");
    m.def("backward", &backward, "backward through the C ABI", pybind11::arg("state"), pybind11::arg("grad_color"), pybind11::arg("grad_depth"),
          pybind11::arg("grad_alpha"), pybind11::arg("accumulate"), pybind11::arg("out"), pybind11::arg("xyz_gradient_accum"),
          pybind11::arg("denom"), pybind11::arg("max_radii2D"), pybind11::arg("push_addr") = 0);
    m.def("abi_version", []() { return dgr_abi_version(); });
    m.def("set_hint", [](int dev, int64_t P, int H, int W, int64_t cap, bool big) {
        std::lock_guard<std::mutex> lk(g_mu); g_hints[hint_key(dev, P, H, W)] = Hint{cap, big}; });
    m.def("get_hint", [](int dev, int64_t P, int H, int W) -> pybind11::object {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_hints.find(hint_key(dev, P, H, W));
        if (it == g_hints.end()) return pybind11::none();
        return pybind11::make_tuple(it->second.cap, it->second.big); });
}
