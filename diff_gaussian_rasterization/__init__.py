"""Drop-in module name of the reference's rasteriser dependency.

VTGaussian-SLAM does `from diff_gaussian_rasterization import GaussianRasterizer as Renderer`
(reference src/vtgaussian_slam.py:38, utils/eval_helpers.py:17) and
`from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera`
(utils/recon_helpers.py:2).  Putting this repo on PYTHONPATH makes those imports resolve to
the B200-native implementation in vtgaussian_slam_b200.
"""
from vtgaussian_slam_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    set_radius_sigma_mult,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "set_radius_sigma_mult"]
