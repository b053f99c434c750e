"""Import alias so reference code runs unchanged:

    from depth_diff_gaussian_rasterization_min import GaussianRasterizationSettings, GaussianRasterizer
    (gaussian_renderer/__init__.py:14 of LucidDreamer)

resolves to the B200-native implementation in luciddreamer_b200.rasterizer.
"""
from luciddreamer_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                          _C, _RasterizeGaussians, rasterize_gaussians)
