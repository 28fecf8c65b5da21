"""View-sharded data parallelism for the rasterizer path.

The reference loops over views sequentially and sums their losses (/root/reference/main.py:219-255); each view's
rasterisation is independent given the Gaussian parameters.  Here the Gaussians are replicated on every GPU, the views
of one iteration are split across ranks (one process per GPU), every rank accumulates the per-Gaussian gradients of its
local views into ONE flat float32 buffer — the backward kernels add straight into it (DgrGaussianGrads.accumulate), so
there is no pack step — and a single NCCL all-reduce(sum) over NVLink makes the buffer identical on all ranks.

Flat layout (floats): [means3D 3P | shs 3MP | opacities P | scales 3P | rotations 4P | means2D 3P].
"""
from typing import List, Optional, Sequence

import torch

from . import rasterizer as _r


class FlatGrads:
    """One contiguous gradient buffer with typed views for every op input."""

    def __init__(self, P: int, M: int, device, with_means2D: bool = True):
        self.P, self.M = P, M
        sizes = [("means3D", 3 * P, (P, 3)), ("shs", 3 * M * P, (P, M, 3)), ("opacities", P, (P, 1)),
                 ("scales", 3 * P, (P, 3)), ("rotations", 4 * P, (P, 4))]
        if with_means2D:
            sizes.append(("means2D", 3 * P, (P, 3)))
        total = sum(n for _, n, _ in sizes)
        self.flat = torch.zeros((total,), dtype=torch.float32, device=device)
        self.views = {}
        o = 0
        for name, n, shape in sizes:
            self.views[name] = self.flat[o:o + n].view(shape)
            o += n
        if not with_means2D:
            self.views["means2D"] = None

    def nbytes(self):
        return self.flat.numel() * 4


class ViewShardedRasterizer:
    """Forward + backward of this rank's views with gradient accumulation, then one all-reduce.

    params: dict(means3D, shs, opacities, scales, rotations) of CUDA float32 tensors (replicated on every rank).
    """

    def __init__(self, P: int, M: int, device, process_group=None):
        self.device = torch.device(device)
        self.grads = FlatGrads(P, M, self.device)
        self.pg = process_group

    def render_views(self, params: dict, settings: Sequence[_r.GaussianRasterizationSettings],
                     upstream: Sequence[tuple], keep_images: bool = False):
        """upstream[i] = (dL_dcolor [3,H,W] or None, dL_ddepth or None, dL_dalpha or None) for local view i."""
        g = self.grads.views
        images = []
        for i, (rs, up) in enumerate(zip(settings, upstream)):
            color, radii, depth, alpha, state = _r.forward_impl(
                rs, params["means3D"], params["shs"], None, params["opacities"], params["scales"], params["rotations"], None)
            _r.backward_impl(state, up[0], up[1], up[2], g["means3D"], g["means2D"], g["shs"], None, g["opacities"],
                             g["scales"], g["rotations"], None, accumulate=(i > 0))
            if keep_images:
                images.append((color, radii, depth, alpha))
        return images

    def all_reduce(self):
        """Sum the flat gradient over ranks (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)
        return self.grads.flat


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition of view indices (rank r gets views [r*V/G, (r+1)*V/G))."""
    per = num_views // world
    rem = num_views % world
    start = rank * per + min(rank, rem)
    return list(range(start, start + per + (1 if rank < rem else 0)))
