"""Host-side mirror of simple_knn._C.distCUDA2 over the C ABI entry dgr_dist_cuda2."""
import ctypes

import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise RuntimeError("distCUDA2 needs a CUDA tensor: this library has no CPU path")
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must have dimensions (num_points, 3)")
    pts = points.detach().float().contiguous()
    P = pts.shape[0]
    lib = _lib.load()
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)          # reference: torch::full({P}, 0.0)
    if P == 0:
        return out
    scratch = torch.empty((lib.dgr_knn_scratch_bytes(P),), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream)
        _lib.check(lib.dgr_dist_cuda2(P, ctypes.c_void_p(pts.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                      ctypes.c_void_p(scratch.data_ptr()), st))
    return out
