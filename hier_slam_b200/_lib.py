"""ctypes binding of libhsraster.so (the C ABI declared in include/hs_raster.h).

There is deliberately NO fallback: if the shared library is missing or fails to load, importing the
rasterizer raises.  Build it with ``python -m hier_slam_b200.build`` (or ``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_size_t, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HS_LIB", os.path.join(HERE, "libhsraster.so"))

# every symbol include/hs_raster.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "hs_abi_version", "hs_last_error", "hs_supports_semantic_channels", "hs_geom_state_bytes", "hs_geom_state_bytes_rows", "hs_forward_readback",
    "hs_image_state_bytes", "hs_binning_state_bytes", "hs_image_state_info_offset", "hs_forward_geometry", "hs_forward_render",
    "hs_backward", "hs_mark_visible", "hs_masked_l1", "hs_hier_cross_entropy", "hs_leaf_cross_entropy", "hs_leaf_cross_entropy_tc", "hs_leaf_ce_workspace_bytes", "hs_leaf_tc_debug", "hs_allreduce_sum", "hs_l1_ssim", "hs_adam_step", "hs_adam_step_device", "hs_transform_points", "hs_tracking_loss", "hs_pose_step", "hs_keyframe_overlap",
    "hs_compact_scratch_bytes", "hs_compact_plan", "hs_compact_gather",
    "hs_geom_state_layout", "hs_image_state_layout",
    "hs_binning_state_layout", "hs_profile_enable", "hs_profile_read", "hs_kernel_launch_count",
    "hs_library_call_count",
]
STAGES = ["preprocess", "scan", "duplicate", "sort", "ranges", "blend_fwd", "blend_bwd", "geom_bwd"]

HS_SEM_ALPHA_EXACT = 1
HS_NO_CULL = 2
HS_BWD_SIMT = 4
HS_SORT_GLOBAL = 32
HS_REUSE_BINNING = 128
HS_DEFER_READBACK = 256
HS_ASYNC_BINNING = 64


class HsCamera(Structure):
    _fields_ = [
        ("image_height", c_int), ("image_width", c_int), ("tanfovx", c_float), ("tanfovy", c_float),
        ("scale_modifier", c_float), ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("bg", c_void_p),
        ("campos", c_void_p), ("prefiltered", c_int), ("debug", c_int),
    ]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run `python -m hier_slam_b200.build`. "
            "There is no CPU / PyTorch fallback for the rasterizer.")
    lib = ctypes.CDLL(LIB_PATH)
    vp = c_void_p
    lib.hs_abi_version.restype = c_int
    lib.hs_last_error.restype = c_char_p
    lib.hs_supports_semantic_channels.argtypes = [c_int]
    lib.hs_supports_semantic_channels.restype = c_int
    lib.hs_geom_state_bytes.argtypes = [c_int]
    lib.hs_geom_state_bytes.restype = c_size_t
    lib.hs_geom_state_bytes_rows.argtypes = [c_int, c_int]
    lib.hs_geom_state_bytes_rows.restype = c_size_t
    lib.hs_forward_readback.argtypes = [POINTER(c_int)]
    lib.hs_forward_readback.restype = c_int
    lib.hs_image_state_bytes.argtypes = [c_int, c_int]
    lib.hs_image_state_bytes.restype = c_size_t
    lib.hs_binning_state_bytes.argtypes = [c_int]
    lib.hs_binning_state_bytes.restype = c_size_t
    lib.hs_image_state_info_offset.argtypes = [c_int, c_int]
    lib.hs_image_state_info_offset.restype = c_size_t
    lib.hs_forward_geometry.argtypes = [POINTER(HsCamera), c_int, vp, vp, vp, vp, vp, vp, c_int, c_int, vp, vp,
                                        c_size_t, vp, c_size_t, c_int, POINTER(c_int), POINTER(c_int), vp]
    lib.hs_forward_geometry.restype = c_int
    lib.hs_forward_render.argtypes = [POINTER(HsCamera), c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_size_t, vp, c_size_t, vp,
                                      c_size_t, vp, vp, vp, vp, vp, vp, c_int, vp]
    lib.hs_forward_render.restype = c_int
    lib.hs_backward.argtypes = ([POINTER(HsCamera), c_int, c_int, c_int] + [vp] * 8 + [c_int, c_int] + [vp] * 21 +
                                [c_int, vp])
    lib.hs_backward.restype = c_int
    lib.hs_masked_l1.argtypes = [vp, vp, vp, c_int, c_size_t, vp, vp, vp]
    lib.hs_masked_l1.restype = c_int
    lib.hs_hier_cross_entropy.argtypes = [vp, vp, c_int, POINTER(c_int), POINTER(c_float), c_size_t, vp, vp, vp]
    lib.hs_hier_cross_entropy.restype = c_int
    lib.hs_leaf_cross_entropy.argtypes = [vp, vp, vp, vp, c_int, c_int, c_size_t, c_float, vp, vp, vp, c_int, vp, vp, vp]
    lib.hs_leaf_cross_entropy.restype = c_int
    lib.hs_leaf_cross_entropy_tc.argtypes = [vp, vp, vp, vp, c_int, c_int, c_size_t, c_float, vp, vp, vp, c_int, vp, vp, vp,
                                             c_size_t, vp]
    lib.hs_leaf_cross_entropy_tc.restype = c_int
    lib.hs_leaf_ce_workspace_bytes.argtypes = [c_int, c_int]
    lib.hs_leaf_ce_workspace_bytes.restype = c_size_t
    lib.hs_allreduce_sum.argtypes = [vp, vp, vp, c_int, c_int, c_size_t, ctypes.c_uint, c_int, vp]
    lib.hs_allreduce_sum.restype = c_int
    lib.hs_leaf_tc_debug.argtypes = [vp]
    lib.hs_leaf_tc_debug.restype = None
    lib.hs_l1_ssim.argtypes = [vp, vp, c_int, c_int, c_int, POINTER(c_float), c_float, c_float, vp, vp, vp, vp]
    lib.hs_l1_ssim.restype = c_int
    lib.hs_adam_step.argtypes = [vp, vp, vp, vp, c_size_t, c_int, POINTER(c_ulonglong), POINTER(c_double), c_double, c_double,
                                 c_double, c_int, vp]
    lib.hs_adam_step.restype = c_int
    lib.hs_adam_step_device.argtypes = [vp, vp, vp, vp, c_size_t, c_int, POINTER(c_ulonglong), POINTER(c_double), c_double,
                                        c_double, c_double, vp, vp, vp]
    lib.hs_adam_step_device.restype = c_int
    lib.hs_transform_points.argtypes = [vp, vp, c_int, vp, vp]
    lib.hs_transform_points.restype = c_int
    lib.hs_tracking_loss.argtypes = [vp, vp, vp, vp, vp, c_size_t, c_float, c_int, c_float, c_float, vp, vp, vp, vp]
    lib.hs_tracking_loss.restype = c_int
    lib.hs_pose_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_float, c_float, c_float, c_float, c_float, c_int, vp]
    lib.hs_pose_step.restype = c_int
    lib.hs_keyframe_overlap.argtypes = [vp, c_int, vp, c_int, c_float, c_float, c_float, c_float, c_int, c_int, c_int, vp, vp]
    lib.hs_keyframe_overlap.restype = c_int
    lib.hs_compact_scratch_bytes.argtypes = [c_int]
    lib.hs_compact_scratch_bytes.restype = c_size_t
    lib.hs_compact_plan.argtypes = [vp, c_int, vp, vp]
    lib.hs_compact_plan.restype = c_int
    lib.hs_compact_gather.argtypes = [vp, vp, vp, c_int, c_int, c_int, POINTER(c_ulonglong), POINTER(c_ulonglong),
                                      POINTER(c_int), vp]
    lib.hs_compact_gather.restype = c_int
    lib.hs_mark_visible.argtypes = [c_int, vp, vp, vp, vp, vp]
    lib.hs_mark_visible.restype = c_int
    lib.hs_geom_state_layout.argtypes = [c_int, POINTER(c_size_t)]
    lib.hs_geom_state_layout.restype = c_int
    lib.hs_image_state_layout.argtypes = [c_int, c_int, POINTER(c_size_t)]
    lib.hs_image_state_layout.restype = c_int
    lib.hs_binning_state_layout.argtypes = [c_int, POINTER(c_size_t)]
    lib.hs_binning_state_layout.restype = c_int
    lib.hs_profile_enable.argtypes = [c_int]
    lib.hs_profile_read.argtypes = [POINTER(c_float), POINTER(c_int)]
    lib.hs_kernel_launch_count.restype = ctypes.c_longlong
    lib.hs_library_call_count.restype = ctypes.c_longlong
    if lib.hs_abi_version() != 5:
        raise ImportError("libhsraster.so ABI version mismatch; rebuild with `python -m hier_slam_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hs_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")


def profile_read() -> dict:
    """{stage: (total_ms, launches)} since the previous read (needs hs_profile_enable(1) before the calls)."""
    lib = load()
    ms = (c_float * 8)()
    cnt = (c_int * 8)()
    check(lib.hs_profile_read(ms, cnt), "hs_profile_read")
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(STAGES) if cnt[i] > 0}
