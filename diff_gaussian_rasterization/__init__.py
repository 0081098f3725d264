"""Drop-in module name of the reference extension (`import diff_gaussian_rasterization`, used at
scripts/hierslam.py:53-54, utils/recon_helpers.py:2, utils/eval_helpers.py:21-22 of the reference).

Put the repository root on PYTHONPATH (instead of installing the reference's extension) and Hier-SLAM's
tracking / mapping code runs against the sm_100a kernels of hier_slam_b200 unchanged."""
from hier_slam_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    GaussianRasterizer_semantic,
    _RasterizeGaussians,
    _RasterizeGaussians_semantic,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
    rasterize_gaussians_semantic,
)
from hier_slam_b200 import _C  # noqa: F401
