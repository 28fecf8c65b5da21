"""ctypes front-end of the C oracle (oracle/dgr_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never by the product package.  PARITY UNPINNED (see dgr_oracle.c header).

The call mirrors the op the reference invokes at /root/reference/gs_renderer.py:745-809:
settings (image size, tan fov, bg, scale_modifier, viewmatrix, projmatrix, sh_degree, campos) + the eight
per-Gaussian inputs -> (color[3,H,W], radii[P] int32, depth[1,H,W], alpha[1,H,W]) and, for the backward,
gradients w.r.t. means3D / means2D / shs / colors_precomp / opacities / scales / rotations / cov3D_precomp.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdgr_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/dgr_oracle.c -> oracle/_build/libdgr_oracle.so (gcc, OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("dgr_oracle.c", "dgr_oracle_impl.h")]
    src.append(os.path.join(_HERE, "..", "include", "dgr_constants.h"))
    if not force and os.path.exists(_LIB_PATH):
        if all(not os.path.exists(s) or os.path.getmtime(s) <= os.path.getmtime(_LIB_PATH) for s in src):
            return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.dgr_oracle_num_threads.restype = ctypes.c_int
        for suf in ("f32", "f64"):
            getattr(_lib, "dgr_oracle_forward_" + suf).restype = ctypes.c_void_p
    return _lib


def num_threads():
    return lib().dgr_oracle_num_threads()


def set_threads(n):
    lib().dgr_oracle_set_threads(int(n))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleResult:
    """Forward outputs + the opaque context the backward needs."""

    def __init__(self, ctx, suf, dt, P, M, H, W, color, depth, alpha, radii):
        self._ctx, self._suf, self._dt = ctx, suf, dt
        self.P, self.M, self.H, self.W = P, M, H, W
        self.color, self.depth, self.alpha, self.radii = color, depth, alpha, radii

    def flags(self):
        """(ambig_px[H,W] bool-ish u8, ambig_g[P] u8 bitmask, N_inst)."""
        px = np.zeros((self.H, self.W), np.uint8)
        g = np.zeros((max(self.P, 1),), np.uint8)
        n = ctypes.c_ulonglong(0)
        getattr(lib(), "dgr_oracle_get_flags_" + self._suf)(ctypes.c_void_p(self._ctx), _ptr(px), _ptr(g), ctypes.byref(n))
        return px, g[: self.P], int(n.value)

    def tie_slack(self):
        """[H,W] float32: how far resolving the pixel's float32 depth-key ties the other way can move its colour."""
        out = np.zeros((self.H, self.W), np.float32)
        getattr(lib(), "dgr_oracle_get_tie_slack_" + self._suf)(ctypes.c_void_p(self._ctx), _ptr(out))
        return out

    def state(self):
        P = max(self.P, 1)
        px, py, dep = (np.zeros(P, self._dt) for _ in range(3))
        conic, rgb = np.zeros((P, 3), self._dt), np.zeros((P, 3), self._dt)
        rect = np.zeros((P, 4), np.int32)
        getattr(lib(), "dgr_oracle_get_state_" + self._suf)(
            ctypes.c_void_p(self._ctx), _ptr(px), _ptr(py), _ptr(dep), _ptr(conic), _ptr(rgb), _ptr(rect))
        s = slice(0, self.P)
        return dict(px=px[s], py=py[s], depth=dep[s], conic=conic[s], rgb=rgb[s], rect=rect[s])

    def backward(self, dL_dcolor=None, dL_ddepth=None, dL_dalpha=None, want_moments=False):
        dt, P, M = self._dt, self.P, self.M
        cv = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dt))
        gC, gD, gA = cv(dL_dcolor), cv(dL_ddepth), cv(dL_dalpha)
        Pn = max(P, 1)
        out = dict(
            means3D=np.zeros((Pn, 3), dt), means2D=np.zeros((Pn, 3), dt), shs=np.zeros((Pn, max(M, 1), 3), dt),
            colors_precomp=np.zeros((Pn, 3), dt), opacities=np.zeros((Pn, 1), dt), scales=np.zeros((Pn, 3), dt),
            rotations=np.zeros((Pn, 4), dt), cov3D_precomp=np.zeros((Pn, 6), dt))
        moments = np.zeros((Pn, 10), np.float64) if want_moments else None
        getattr(lib(), "dgr_oracle_backward_" + self._suf)(
            ctypes.c_void_p(self._ctx), _ptr(gC), _ptr(gD), _ptr(gA),
            _ptr(out["means3D"]), _ptr(out["means2D"]), _ptr(out["shs"]), _ptr(out["colors_precomp"]),
            _ptr(out["opacities"]), _ptr(out["scales"]), _ptr(out["rotations"]), _ptr(out["cov3D_precomp"]),
            _ptr(moments))
        res = {k: v[:P] for k, v in out.items()}
        if want_moments:
            res["_moments"] = moments[:P]
        return res

    def close(self):
        if self._ctx:
            getattr(lib(), "dgr_oracle_free_" + self._suf)(ctypes.c_void_p(self._ctx))
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forward(*, image_height, image_width, tanfovx, tanfovy, bg, scale_modifier, viewmatrix, projmatrix,
            sh_degree, campos, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, prefiltered=False, dtype=np.float64, eps=2e-5):
    """Run the oracle forward.  All array arguments are numpy (any float dtype); see module docstring."""
    dt = np.dtype(dtype)
    suf = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[dt]
    cv = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dt))
    means3D = cv(means3D)
    P = means3D.shape[0]
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(cv, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    M = 0 if shs is None else shs.shape[1]
    H, W = int(image_height), int(image_width)
    color = np.zeros((3, H, W), dt); depth = np.zeros((1, H, W), dt); alpha = np.zeros((1, H, W), dt)
    radii = np.zeros((max(P, 1),), np.int32)
    ctx = getattr(lib(), "dgr_oracle_forward_" + suf)(
        ctypes.c_int(P), ctypes.c_int(M), ctypes.c_int(int(sh_degree)), ctypes.c_int(H), ctypes.c_int(W),
        ctypes.c_double(tanfovx), ctypes.c_double(tanfovy), ctypes.c_double(scale_modifier),
        _ptr(cv(bg)), _ptr(cv(viewmatrix)), _ptr(cv(projmatrix)), _ptr(cv(campos)),
        _ptr(means3D), _ptr(shs), _ptr(colors_precomp), _ptr(cv(opacities).reshape(-1)),
        _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp),
        ctypes.c_int(int(prefiltered)), ctypes.c_double(eps),
        _ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii))
    if not ctx:
        raise RuntimeError("dgr_oracle_forward rejected its arguments")
    return OracleResult(ctx, suf, dt, P, M, H, W, color, depth, alpha, radii[:P])
