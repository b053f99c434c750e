"""B200-native differentiable 3D Gaussian Splatting rasterizer (drop-in for LucidDreamer's
depth_diff_gaussian_rasterization_min).  Importing the operator API loads the CUDA library; there is no fallback."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, _C, _RasterizeGaussians,  # noqa: F401
                         rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "_RasterizeGaussians", "_C"]
