"""ctypes binding of libenvgs_b200.so (the C ABI in include/envgs_b200.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing the operator
modules raises.  The oracle under `oracle/` is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ENVGS_B200_LIB") or os.path.join(_HERE, "libenvgs_b200.so")   # override: kernel A/B experiments only

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_fp = C.c_void_p  # device pointers travel as plain addresses


class RasterArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("C", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("opacities", _fp),
        ("scales", _fp), ("rotations", _fp), ("transMat_precomp", _fp), ("viewmatrix", _fp), ("projmatrix", _fp),
        ("campos", _fp),
        ("instance_capacity", C.c_int32), ("instance_count_out", _fp),
    ]


class RasterOut(C.Structure):
    _fields_ = [("out_color", _fp), ("out_others", _fp), ("out_weight", _fp), ("radii", _fp)]


class RasterGrads(C.Structure):
    _fields_ = [
        ("dL_dout_color", _fp), ("dL_dout_others", _fp), ("dL_dmeans2D", _fp), ("dL_dcolors", _fp),
        ("dL_dopacity", _fp), ("dL_dmeans3D", _fp), ("dL_dtransMat", _fp), ("dL_dsh", _fp), ("dL_dscales", _fp),
        ("dL_drotations", _fp), ("dL_dnormal", _fp),
    ]


class TraceArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("training", C.c_int32), ("start_from_first", C.c_int32), ("max_trace_depth", C.c_int32),
        ("specular_threshold", C.c_float), ("scale_modifier", C.c_float),
        ("ray_o", _fp), ("ray_d", _fp), ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp),
        ("others_precomp", _fp), ("opacities", _fp), ("scales", _fp), ("rotations", _fp), ("viewmatrix", _fp),
        ("projmatrix", _fp), ("campos", _fp),
    ]


class GlueCamera(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("K", _fp), ("R", _fp), ("campos", _fp)]


class TraceOut(C.Structure):
    _fields_ = [("out_rgb", _fp), ("out_dpt", _fp), ("out_acc", _fp), ("out_norm", _fp), ("out_dist", _fp),
                ("out_aux", _fp), ("mid_val", _fp), ("a_weights", _fp), ("hit_cache", _fp)]


class TraceGrads(C.Structure):
    _fields_ = [
        ("out_rgb", _fp), ("out_dpt", _fp), ("out_acc", _fp), ("out_norm", _fp), ("out_dist", _fp), ("out_aux", _fp),
        ("mid_val", _fp),
        ("dL_drgb", _fp), ("dL_ddpt", _fp), ("dL_dacc", _fp), ("dL_dnorm", _fp), ("dL_ddist", _fp), ("dL_daux", _fp),
        ("dL_dray_o", _fp), ("dL_dray_d", _fp), ("dL_dmeans3D", _fp), ("dL_dshs", _fp), ("dL_dcolors", _fp),
        ("dL_dothers", _fp), ("dL_dopacities", _fp), ("dL_dscales", _fp), ("dL_drotations", _fp), ("hit_cache", _fp),
    ]


class DensifyParams(C.Structure):
    _fields_ = [("P", C.c_int32), ("n_split", C.c_int32), ("split_ratio", C.c_float), ("grad_threshold", C.c_float),
                ("size_threshold", C.c_float), ("split_screen_threshold", C.c_float), ("min_opacity", C.c_float),
                ("min_gradient", C.c_float)]


class RowArray(C.Structure):
    _fields_ = [("src", _fp), ("dst", _fp), ("width", C.c_int32), ("role", C.c_int32)]


ROW_PLAIN, ROW_XYZ, ROW_SCALING, ROW_OPT_STATE, ROW_STAT_GRAD, ROW_STAT_DENOM, ROW_STAT_RADII, ROW_STAT_WEIGHT = 0, 1, 2, 4, 5, 6, 7, 8


# every symbol include/envgs_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "envgs_last_error", "envgs_version", "envgs_raster_forward", "envgs_raster_backward", "envgs_mark_visible",
    "envgs_raster_geom_bytes", "envgs_raster_img_bytes", "envgs_raster_binning_bytes", "envgs_tracer_create",
    "envgs_tracer_destroy", "envgs_tracer_build", "envgs_tracer_bvh_bytes", "envgs_trace_forward",
    "envgs_trace_backward", "envgs_launch_count", "envgs_profile_enable", "envgs_profile_report", "envgs_debug_trace_stats", "envgs_trace_cache_bytes", "envgs_trace_cache_bytes2", "envgs_sh_features_forward", "envgs_sh_features_backward", "envgs_sh_features_backward_factored", "envgs_sh_adam_factored",
    "envgs_glue_reflect_forward", "envgs_glue_reflect_backward", "envgs_glue_blend_loss_forward", "envgs_glue_blend_loss_backward",
    "envgs_adam_step", "envgs_glue_quad_vertices",
    "envgs_densify_stats", "envgs_densify_plan_scratch_bytes", "envgs_densify_plan", "envgs_densify_apply", "envgs_rows_gather",
]

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m envgs_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.envgs_last_error.restype = C.c_char_p
    L.envgs_version.restype = C.c_int
    L.envgs_launch_count.restype = C.c_uint64
    L.envgs_raster_forward.restype = C.c_int
    L.envgs_raster_forward.argtypes = [C.POINTER(RasterArgs), C.POINTER(RasterOut), ALLOC_FN, ALLOC_FN, ALLOC_FN,
                                       C.c_void_p, C.c_void_p]
    L.envgs_raster_backward.restype = C.c_int
    L.envgs_raster_backward.argtypes = [C.POINTER(RasterArgs), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(RasterGrads), C.c_void_p]
    L.envgs_mark_visible.restype = C.c_int
    L.envgs_mark_visible.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("envgs_raster_geom_bytes", "envgs_raster_binning_bytes", "envgs_tracer_bvh_bytes"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_int32]
    L.envgs_raster_img_bytes.restype = C.c_size_t
    L.envgs_raster_img_bytes.argtypes = [C.c_int32, C.c_int32]
    L.envgs_tracer_create.restype = C.c_void_p
    L.envgs_tracer_create.argtypes = []
    L.envgs_tracer_destroy.restype = None
    L.envgs_tracer_destroy.argtypes = [C.c_void_p]
    L.envgs_tracer_build.restype = C.c_int
    L.envgs_tracer_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, ALLOC_FN, C.c_void_p, C.c_void_p]
    L.envgs_trace_forward.restype = C.c_int
    L.envgs_trace_forward.argtypes = [C.c_void_p, C.POINTER(TraceArgs), C.POINTER(TraceOut), C.c_void_p]
    L.envgs_trace_backward.restype = C.c_int
    L.envgs_trace_backward.argtypes = [C.c_void_p, C.POINTER(TraceArgs), C.POINTER(TraceGrads), C.c_void_p]
    L.envgs_profile_enable.restype = None
    L.envgs_profile_enable.argtypes = [C.c_int]
    L.envgs_profile_report.restype = C.c_size_t
    L.envgs_profile_report.argtypes = [C.c_char_p, C.c_size_t]
    L.envgs_trace_cache_bytes.restype = C.c_size_t
    L.envgs_trace_cache_bytes.argtypes = [C.c_int32, C.c_int32]
    L.envgs_trace_cache_bytes2.restype = C.c_size_t
    L.envgs_trace_cache_bytes2.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.envgs_sh_features_forward.restype = C.c_int
    L.envgs_sh_features_forward.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 6
    L.envgs_sh_features_backward.restype = C.c_int
    L.envgs_sh_features_backward.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 8
    L.envgs_sh_features_backward_factored.restype = C.c_int
    L.envgs_sh_features_backward_factored.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 8
    L.envgs_sh_adam_factored.restype = C.c_int
    L.envgs_sh_adam_factored.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 6 + [C.c_int32] + [C.c_float] * 5 + [C.c_void_p]
    L.envgs_glue_reflect_forward.restype = C.c_int
    L.envgs_glue_reflect_forward.argtypes = [C.POINTER(GlueCamera)] + [C.c_void_p] * 4
    L.envgs_glue_reflect_backward.restype = C.c_int
    L.envgs_glue_reflect_backward.argtypes = [C.POINTER(GlueCamera)] + [C.c_void_p] * 5
    L.envgs_glue_blend_loss_forward.restype = C.c_int
    L.envgs_glue_blend_loss_forward.argtypes = [C.POINTER(GlueCamera)] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 3
    L.envgs_glue_blend_loss_backward.restype = C.c_int
    L.envgs_glue_blend_loss_backward.argtypes = [C.POINTER(GlueCamera)] + [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 6
    L.envgs_glue_quad_vertices.restype = C.c_int
    L.envgs_glue_quad_vertices.argtypes = [C.c_int32] + [C.c_void_p] * 5
    L.envgs_adam_step.restype = C.c_int
    L.envgs_adam_step.argtypes = [C.c_uint64] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_float] * 5 + [C.c_int32, C.c_void_p]
    L.envgs_debug_trace_stats.restype = C.c_int
    L.envgs_debug_trace_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
    L.envgs_densify_stats.restype = C.c_int
    L.envgs_densify_stats.argtypes = [C.c_int32] + [C.c_void_p] * 8
    L.envgs_densify_plan_scratch_bytes.restype = C.c_size_t
    L.envgs_densify_plan_scratch_bytes.argtypes = [C.c_int32]
    L.envgs_densify_plan.restype = C.c_int
    L.envgs_densify_plan.argtypes = [C.POINTER(DensifyParams)] + [C.c_void_p] * 7 + [C.POINTER(C.c_int32), C.c_void_p, C.c_size_t, C.c_void_p]
    L.envgs_densify_apply.restype = C.c_int
    L.envgs_densify_apply.argtypes = [C.POINTER(DensifyParams), C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(RowArray)] + \
        [C.c_void_p] * 6
    L.envgs_rows_gather.restype = C.c_int
    L.envgs_rows_gather.argtypes = [C.c_int64, C.c_void_p, C.c_int32, C.POINTER(RowArray), C.c_void_p]
    _lib = L
    return L


def check(rc: int, what: str) -> int:
    if rc < 0:
        msg = lib().envgs_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
    return rc


def ptr(t) -> int | None:
    """Device address of a torch tensor, or None (NULL) for an empty / missing one."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().envgs_launch_count())


def profile_enable(on: bool) -> None:
    lib().envgs_profile_enable(int(bool(on)))


def profile_report() -> dict:
    """{kernel name: (launches, total device ms)} since the last report (CUDA events on the launching stream)."""
    buf = C.create_string_buffer(1 << 16)
    lib().envgs_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
