"""dreamgaussian_b200 — B200-native differentiable 3D Gaussian-splat rasterizer (hot path of DreamGaussian).

Public surface = the reference's operator surface (see rasterizer.py): GaussianRasterizationSettings, GaussianRasterizer.
Importing this package does not need a GPU; calling the rasterizer does (there is no CPU path).
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
