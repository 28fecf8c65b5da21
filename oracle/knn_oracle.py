"""ctypes front-end of oracle/knn_oracle.c (CPU restatement of distCUDA2) and of oracle/_ref/libsimple_knn_ref.so (the
reference's own simple_knn.cu, compiled unmodified).  TEST INFRASTRUCTURE ONLY — tests/, tools/ and bench legs; never
the product."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libknn_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libsimple_knn_ref.so")
_lib = None


def build():
    src = os.path.join(_HERE, "knn_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libknn_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def build_ref():
    """oracle/_ref from the reference sources (only where /root/reference exists; the GPU box uses the prebuilt file)."""
    if os.path.exists("/root/reference/simple-knn/simple_knn.cu"):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"], stdout=subprocess.DEVNULL)
    return REF_LIB if os.path.exists(REF_LIB) else None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def dist2_f32(points):
    p = np.ascontiguousarray(points, np.float32)
    out = np.empty((p.shape[0],), np.float32)
    _load().knn_oracle_f32(ctypes.c_int(p.shape[0]), p.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    return out


def dist2_f64(points):
    p = np.ascontiguousarray(points, np.float32)
    out = np.empty((p.shape[0],), np.float64)
    _load().knn_oracle_f64(ctypes.c_int(p.shape[0]), p.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    return out


def reference_dist_cuda2(points_cuda):
    """The reference's own CUDA code on a torch CUDA tensor [P,3] float32 (needs a GPU and oracle/_ref)."""
    import torch
    lib = ctypes.CDLL(REF_LIB)
    pts = points_cuda.contiguous().float()
    out = torch.empty((pts.shape[0],), dtype=torch.float32, device=pts.device)
    torch.cuda.synchronize(pts.device)
    with torch.cuda.device(pts.device):
        rc = lib.ref_dist_cuda2(ctypes.c_int(pts.shape[0]), ctypes.c_void_p(pts.data_ptr()), ctypes.c_void_p(out.data_ptr()))
    if rc != 0:
        raise RuntimeError("reference simple_knn failed: cuda error %d" % rc)
    return out
