"""Host-side mirror of `GaussianModel.extract_fields` (/root/reference/gs_renderer.py:218-294) over the
C ABI entry dgr_extract_fields."""
import ctypes

import torch

from . import _lib


def extract_fields(xyz, opacity, scaling, rotation, resolution=128, num_blocks=16, relax_ratio=1.5):
    """Raw model tensors (_xyz, _opacity, _scaling, _rotation) -> (occ [res,res,res] float32, center [3], scale [] ) on the
    device; nothing is synchronised (read `scale.item()` when the host needs it, as extract_mesh does)."""
    for name, t in (("xyz", xyz), ("opacity", opacity), ("scaling", scaling), ("rotation", rotation)):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor: extract_fields has no CPU path" % name)
    if resolution % num_blocks:
        raise ValueError("resolution must be a multiple of num_blocks")
    dev = xyz.device
    x, o, s, r = (t.detach().float().contiguous() for t in (xyz, opacity, scaling, rotation))
    P = x.shape[0]
    if o.numel() != P or s.numel() != 3 * P or r.numel() != 4 * P:
        raise ValueError("opacity / scaling / rotation must have num_points rows")
    lib = _lib.load()
    occ = torch.empty((resolution,) * 3, dtype=torch.float32, device=dev)
    cs = torch.zeros((4,), dtype=torch.float32, device=dev)
    scratch = torch.empty((lib.dgr_fields_scratch_bytes(P, num_blocks),), dtype=torch.uint8, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.dgr_extract_fields(P, p(x), p(o), p(s), p(r), int(resolution), int(num_blocks), ctypes.c_float(relax_ratio),
                                          p(occ), p(cs), p(scratch), st))
    return occ, cs[:3], cs[3]


def extract_fields_of_model(gaussians, resolution=128, num_blocks=16, relax_ratio=1.5):
    """Same call shape as the reference method: sets gaussians.center / gaussians.scale (:237-238) and returns occ."""
    occ, center, scale = extract_fields(gaussians._xyz, gaussians._opacity, gaussians._scaling, gaussians._rotation,
                                        resolution, num_blocks, relax_ratio)
    gaussians.center = center
    gaussians.scale = float(scale.item())
    return occ
