"""dreamgaussian_b200 — B200-native differentiable 3D Gaussian-splat rasterizer (hot path of DreamGaussian).

Public surface = the reference's operator surface (see rasterizer.py): GaussianRasterizationSettings, GaussianRasterizer.
Importing this package does not need a GPU; calling the rasterizer does (there is no CPU path).

The operator names resolve lazily (PEP 562): `import dreamgaussian_b200.scene` (pure numpy; what bench.py's CPU reference
arm needs) does not load libdgr_b200.so / dgr_torch_host.so into the process.
"""
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]


def __getattr__(name):
    if name in __all__:
        from . import rasterizer
        return getattr(rasterizer, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
