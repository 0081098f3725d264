"""hier_slam_b200 — Blackwell-native (sm_100a) differentiable Gaussian rasterizer for Hier-SLAM's render hot path.

Only the hot path lives here: `csrc/` (CUDA kernels + the C ABI of include/hs_raster.h),
`_lib` (ctypes binding, fails loudly when the library is missing), `_C` (the reference's five native entry points),
`rasterizer` (the reference's Python API), `mapping` (keyframe-parallel multi-GPU mapping) and `scene` (seeded
synthetic scenes for tests and benchmarks)."""
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    GaussianRasterizer_semantic,
    rasterize_gaussians,
    rasterize_gaussians_semantic,
)

__version__ = "0.1.0"
