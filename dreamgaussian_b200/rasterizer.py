"""Host-side mirror of the reference's rasterizer operator: ``GaussianRasterizationSettings`` / ``GaussianRasterizer``.

Drop-in for the import at /root/reference/gs_renderer.py:10-13 and the call at gs_renderer.py:745-809:

    rasterizer = GaussianRasterizer(raster_settings=GaussianRasterizationSettings(...12 fields...))
    color, radii, depth, alpha = rasterizer(means3D=, means2D=, shs=, colors_precomp=, opacities=,
                                            scales=, rotations=, cov3D_precomp=)

Same names, argument meaning, return shapes/dtypes (color [3,H,W], radii int32 [P], depth [1,H,W], alpha [1,H,W]) and
error behaviour (the two "Please provide ..." exceptions) as the un-vendored package the reference imports [EXT].
All arithmetic runs in libdgr_b200.so (hand-written sm_100a CUDA) through the C ABI of include/dgr_b200.h; torch is
used for device memory, the current stream and autograd plumbing only.  There is no CPU path: non-CUDA tensors raise.
"""
import ctypes
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib


def _load_fast():
    """The compiled host layer (csrc/dgr_torch.cpp -> lib/dgr_torch_host.so): same C-ABI calls, ~10x less host time per
    call than ctypes.  Optional: without it (or with DGR_NO_FASTHOST=1) the ctypes path below does the same work."""
    import importlib.util
    import os
    if os.environ.get("DGR_NO_FASTHOST") == "1":
        return None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "dgr_torch_host.so")
    if not os.path.exists(path):
        return None
    try:
        from . import build as _build
        if not os.path.exists(_build.HOST_HASH) or open(_build.HOST_HASH).read().strip() != _build._host_hash():
            import warnings
            warnings.warn("dgr_torch_host.so is stale (csrc/dgr_torch.cpp or include/dgr_b200.h changed since it was built): "
                          "using the ctypes host layer; run dreamgaussian_b200/build.py", RuntimeWarning)
            return None
        _lib.load()
        spec = importlib.util.spec_from_file_location("dgr_torch_host", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod if mod.abi_version() == _lib.ABI_VERSION else None
    except Exception as e:  # noqa: BLE001  (a torch / compiler mismatch: fall back to ctypes, never to a CPU path)
        import warnings
        warnings.warn("dgr_torch_host.so could not be loaded (%r): using the ctypes host layer" % (e,), RuntimeWarning)
        return None


_FAST = _load_fast()


def set_fast_host(enabled: bool):
    """Switch between the compiled host layer and the ctypes one (tests exercise both)."""
    global _FAST
    _FAST = _load_fast() if enabled else None
    return _FAST is not None


def get_capacity_hint(device_index, P, H, W):
    if _FAST is not None:
        return _FAST.get_hint(device_index, P, H, W)
    return _CAPACITY_HINT.get((device_index, P, H, W))


def set_capacity_hint(device_index, P, H, W, cap, big):
    if _FAST is not None:
        _FAST.set_hint(device_index, P, H, W, cap, big)
    else:
        _CAPACITY_HINT[(device_index, P, H, W)] = (cap, big)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError(
            "%s is on %s: this rasterizer has no CPU path (it runs as sm_100a CUDA kernels in libdgr_b200.so)" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return _aligned16(t.contiguous())


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    """The kernels read rotations / SH rows with 128-bit loads and stage their inputs with bulk TMA: a contiguous view that starts
    off a 16-byte boundary (e.g. ``flat[1:].view(P, 4)``) gets its own allocation, which is aligned."""
    if t.numel() > 0 and t.data_ptr() % 16 != 0:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _opt(t, name):
    if t is None or (isinstance(t, torch.Tensor) and t.numel() == 0 and t.dim() <= 1):
        return None
    return _dev_f32(t, name)


class _Frame:
    """ctypes views of one forward call's settings + inputs (keeps the tensors alive)."""

    def __init__(self, rs: GaussianRasterizationSettings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3D,
                 sh_rest=None, activations=False):
        self.keep = []
        k = self.keep.append
        self.bg = _dev_f32(rs.bg, "bg"); k(self.bg)
        self.view = _dev_f32(rs.viewmatrix, "viewmatrix"); k(self.view)
        self.proj = _dev_f32(rs.projmatrix, "projmatrix"); k(self.proj)
        self.campos = _dev_f32(rs.campos, "campos"); k(self.campos)
        if self.bg.numel() != 3 or self.view.numel() != 16 or self.proj.numel() != 16 or self.campos.numel() != 3:
            raise ValueError("bg/campos must have 3 elements, viewmatrix/projmatrix 16")
        self.settings = _lib.DgrSettings(
            int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier),
            int(rs.sh_degree), int(bool(rs.prefiltered)), int(bool(rs.debug)),
            _ptr(self.bg), _ptr(self.view), _ptr(self.proj), _ptr(self.campos))
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise ValueError("means3D must have dimensions (num_points, 3)")
        P = means3D.shape[0]
        M = 0
        if sh is not None:
            if sh.dim() != 3 or sh.shape[0] != P or sh.shape[2] != 3:
                raise ValueError("shs must have dimensions (num_points, num_coeffs, 3)")
            M = sh.shape[1]
        if activations:
            # raw GaussianModel parameters: sh = _features_dc [P,1,3], sh_rest = _features_rest [P,M-1,3]
            if sh is None or sh.shape[1] != 1:
                raise ValueError("features_dc must have dimensions (num_points, 1, 3)")
            if sh_rest is not None and sh_rest.numel() > 0:
                if sh_rest.dim() != 3 or sh_rest.shape[0] != P or sh_rest.shape[2] != 3:
                    raise ValueError("features_rest must have dimensions (num_points, num_coeffs - 1, 3)")
                M = 1 + sh_rest.shape[1]
            else:
                sh_rest = None
        for t, n, w in ((colors_precomp, "colors_precomp", 3), (scales, "scales", 3), (rotations, "rotations", 4), (cov3D, "cov3D_precomp", 6)):
            if t is not None and (t.numel() != P * w):
                raise ValueError("%s must have dimensions (num_points, %d)" % (n, w))
        if opacities.numel() != P:
            raise ValueError("opacities must have dimensions (num_points, 1)")
        self.P, self.M = P, M
        self.gaussians = _lib.DgrGaussians(P, M, _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(opacities),
                                           _ptr(scales), _ptr(rotations), _ptr(cov3D), _ptr(sh_rest if activations else None),
                                           1 if activations else 0)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ForwardState:
    """Everything one forward leaves behind for its backward (upstream: ctx + geom/binning/img buffers)."""
    __slots__ = ("rs", "frame", "num_rendered", "capacity", "geom", "binning", "image", "radii", "alpha", "tensors", "fast",
                 "stats_applied")


def forward_impl(rs, means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, sh_rest=None, activations=False):
    """Runs both forward stages through the C ABI. Inputs are validated CUDA float32 tensors (or None).
    Returns (color, radii, depth, alpha, ForwardState).

    No host round trip sits between the kernels: the instance buffer is sized from a per-shape capacity hint, every
    kernel of the forward is enqueued, and only then does the host wait for the (early) instance-count event.  If the
    guess was too small, stage 2 is re-run with a large enough buffer (the kernels clip safely to the capacity)."""
    if _FAST is not None:
        color, radii, depth, alpha, cst = _FAST.forward(
            int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier), int(rs.sh_degree),
            bool(rs.prefiltered), bool(rs.debug), rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos, means3D, sh, colors_precomp, opacities,
            scales, rotations, cov3D, sh_rest, bool(activations))
        state = ForwardState()
        state.fast, state.rs, state.num_rendered, state.capacity = cst, rs, cst.num_rendered, cst.capacity
        state.radii, state.alpha = radii, None
        state.tensors = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, sh_rest)
        return color, radii, depth, alpha, state
    lib = _lib.load()
    dev = means3D.device
    with torch.cuda.device(dev):
        fr = _Frame(rs, means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, sh_rest, activations)
        H, W, P = int(rs.image_height), int(rs.image_width), fr.P
        u8 = dict(dtype=torch.uint8, device=dev)
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        alpha = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty((lib.dgr_geom_bytes(P, H, W),), **u8)
        image = torch.empty((lib.dgr_image_bytes(H, W),), **u8)
        n_host, event = _host_sync_objects(dev)
        st = _stream_ptr(dev)
        _lib.check(lib.dgr_forward_preprocess(ctypes.byref(fr.settings), ctypes.byref(fr.gaussians), _ptr(geom), _ptr(image),
                                              _ptr(radii), st))
        out = _lib.DgrImages(_ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii))
        key = (dev.index, P, H, W)
        cap, big = _CAPACITY_HINT.get(key, (max(65536, 16 * P), True))
        rerun = 0
        while True:
            binning = torch.empty((lib.dgr_binning_bytes(cap, H, W),), **u8)
            _lib.check(lib.dgr_forward_render(ctypes.byref(fr.settings), ctypes.byref(fr.gaussians), _ptr(geom), _ptr(binning),
                                              ctypes.c_uint64(cap), _ptr(image), ctypes.byref(out), (1 if big else 0) | rerun,
                                              ctypes.c_void_p(n_host.data_ptr()), ctypes.c_uint64(0), event, st))
            _lib.check(lib.dgr_event_synchronize(event))
            n_inst, n_big = int(n_host[0]), int(n_host[1])
            if n_inst <= cap and (big or n_big == 0):
                break
            cap = max(cap, int(n_inst * 1.25) + 4096)  # a guess was wrong: redo stage 2 (rare)
            big = big or n_big > 0
            rerun = 2                                   # DGR_FLAG_RERUN
        _CAPACITY_HINT[key] = (max(int(n_inst * 1.25) + 4096, 65536), n_big > 0)
    state = ForwardState()
    state.fast = None
    state.rs, state.frame, state.num_rendered, state.capacity = rs, fr, n_inst, cap
    # alpha is a differentiable OUTPUT: storing it here would close the cycle state -> alpha -> grad_fn -> ctx -> state and leave
    # ~100 MB of scratch per call to the cyclic garbage collector (the C ABI's out_alpha argument is unused)
    state.geom, state.binning, state.image, state.radii, state.alpha = geom, binning, image, radii, None
    state.tensors = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, sh_rest)
    return color, radii, depth, alpha, state


_SYNC = {}
_CAPACITY_HINT = {}


def _host_sync_objects(dev):
    """(pinned uint64 for the instance count, cudaEvent handle) per (host thread, device)."""
    import threading
    key = (threading.get_ident(), dev.index)
    t = _SYNC.get(key)
    if t is None:
        ev = _lib.load().dgr_event_create()
        if not ev:
            raise RuntimeError("libdgr_b200: could not create a CUDA event")
        t = (torch.zeros((4,), dtype=torch.int64).pin_memory(), ctypes.c_void_p(ev))
        _SYNC[key] = t
    return t


def backward_impl(state, grad_color, grad_depth, grad_alpha, d_means3D, d_means2D, d_sh, d_col, d_opac, d_scales, d_rot,
                  d_cov, accumulate=False, d_sh_rest=None, densify=None, push=None):
    """Runs the backward through the C ABI, writing (or accumulating) into the given gradient tensors.
    densify = (xyz_gradient_accum, denom, max_radii2D) float32 [P] tensors (any may be None) updated in the same kernel.
    push = a _lib.DgrPeerPush (multi-GPU: rows of Gaussians another rank owns go straight to that rank, include/dgr_b200.h)."""
    push_addr = ctypes.addressof(push) if push is not None else 0
    if state.fast is not None:
        _FAST.backward(state.fast, grad_color, grad_depth, grad_alpha, bool(accumulate),
                       [d_means3D, d_means2D, d_sh, d_col, d_opac, d_scales, d_rot, d_cov, d_sh_rest], *(densify or (None, None, None)),
                       push_addr)
        return
    lib = _lib.load()
    fr = state.frame
    dev = state.radii.device
    with torch.cuda.device(dev):
        gin = _lib.DgrImageGrads(_ptr(grad_color), _ptr(grad_depth), _ptr(grad_alpha))
        gout = _lib.DgrGaussianGrads(_ptr(d_means3D), _ptr(d_means2D), _ptr(d_sh), _ptr(d_col), _ptr(d_opac), _ptr(d_scales),
                                     _ptr(d_rot), _ptr(d_cov), 1 if accumulate else 0, _ptr(d_sh_rest),
                                     *((_ptr(t) for t in densify) if densify is not None else (None, None, None)), push_addr or None)
        _lib.check(lib.dgr_backward(ctypes.byref(fr.settings), ctypes.byref(fr.gaussians), _ptr(state.geom), _ptr(state.binning),
                                    ctypes.c_uint64(state.capacity), _ptr(state.image), _ptr(state.radii), _ptr(state.alpha),
                                    ctypes.byref(gin), ctypes.byref(gout), _stream_ptr(dev)))


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        means3D = _dev_f32(means3D, "means3D")
        opacities = _dev_f32(opacities, "opacities")
        sh, colors_precomp = _opt(sh, "shs"), _opt(colors_precomp, "colors_precomp")
        scales, rotations, cov3D = _opt(scales, "scales"), _opt(rotations, "rotations"), _opt(cov3Ds_precomp, "cov3D_precomp")
        color, radii, depth, alpha, state = forward_impl(raster_settings, means3D, sh, colors_precomp, opacities, scales,
                                                         rotations, cov3D)
        ctx.state = state
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        state = ctx.state
        if state.fast is not None:
            g = _FAST.backward(state.fast, grad_color, grad_depth, grad_alpha, False, [], None, None, None)      # undefined -> None
            return g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], None
        means3D, sh, colors_precomp, opacities, scales, rotations, cov3D, _ = state.tensors
        dev = means3D.device
        P, M = state.frame.P, state.frame.M
        f32 = dict(dtype=torch.float32, device=dev)
        gC = _dev_f32(grad_color, "grad_color") if grad_color is not None else None
        gD = _dev_f32(grad_depth, "grad_depth") if grad_depth is not None else None
        gA = _dev_f32(grad_alpha, "grad_alpha") if grad_alpha is not None else None
        d_means3D = torch.empty((P, 3), **f32)
        d_means2D = torch.empty((P, 3), **f32)
        d_opac = torch.empty((P, 1), **f32)
        d_sh = torch.empty((P, M, 3), **f32) if sh is not None else None
        d_col = torch.empty((P, 3), **f32) if colors_precomp is not None else None
        d_scales = torch.empty((P, 3), **f32) if scales is not None else None
        d_rot = torch.empty((P, 4), **f32) if rotations is not None else None
        d_cov = torch.empty((P, 6), **f32) if cov3D is not None else None
        backward_impl(state, gC, gD, gA, d_means3D, d_means2D, d_sh, d_col, d_opac, d_scales, d_rot, d_cov)
        # ctx.state stays: a retain_graph backward runs again from the same buffers (they go away with the graph node)
        return d_means3D, d_means2D, d_sh, d_col, d_opac, d_scales, d_rot, d_cov, None


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizer(nn.Module):
    """Constructed per render call by the reference (gs_renderer.py:760): construction is trivial."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            lib = _lib.load()
            pos = _dev_f32(positions, "positions")
            view, proj = _dev_f32(rs.viewmatrix, "viewmatrix"), _dev_f32(rs.projmatrix, "projmatrix")
            present = torch.empty((pos.shape[0],), dtype=torch.uint8, device=pos.device)
            with torch.cuda.device(pos.device):
                _lib.check(lib.dgr_mark_visible(pos.shape[0], _ptr(pos), _ptr(view), _ptr(proj), _ptr(present), _stream_ptr(pos.device)))
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings)
