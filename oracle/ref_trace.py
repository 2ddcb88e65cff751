"""ORACLE tooling (checker only): ctypes binding of oracle/_ref/ref_trace/libref_trace_{fwd,bwd}.so — the UNMODIFIED
reference tracer device programs (optix_tracer/forward.cu, backward.cu) compiled against the OptiX emulation header
oracle/optix_shim/optix.h by oracle/build_ref.py.  Never imported by the product path.

`trace_forward` / `trace_backward` take and return what the reference's pybind functions `trace_surfels` /
`trace_surfels_backward` do (/root/reference/submodules/diff-surfel-tracing/trace_surfels.cpp:173-315, :318-471):
outputs are allocated zero-filled here exactly as there, empty tensors become null pointers, `out_dist` is (H,W,3).
"""
from __future__ import annotations

import ctypes
import os

import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_trace")
_FP = ctypes.c_void_p


class _Args(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ("P", "H", "W", "D", "M", "training", "start_from_first", "max_trace_depth")]
                + [("specular_threshold", ctypes.c_float), ("scale_modifier", ctypes.c_float), ("num_triangles", ctypes.c_int)]
                + [(n, _FP) for n in (
                    "vertices", "triangles",
                    "ray_o", "ray_d", "background", "means3D", "shs", "colors_precomp", "others_precomp", "opacities", "scales",
                    "rotations", "viewmatrix", "projmatrix", "campos",
                    "out_rgb", "out_dpt", "out_acc", "out_dist", "out_norm", "out_aux", "mid_val", "a_weights",
                    "dL_drgb", "dL_ddpt", "dL_dacc", "dL_ddist", "dL_dnorm", "dL_daux",
                    "dL_dray_o", "dL_dray_d", "dL_dmeans3D", "dL_dshs", "dL_dcolors", "dL_dothers", "dL_dopacities", "dL_dscales",
                    "dL_drotations", "dL_dtransMat")])


_LIBS: dict = {}
OPTIX_OUT = os.path.join(os.path.dirname(OUT), "ref_trace_optix")
_OPTIX: dict = {}


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, f"libref_trace_{k}.so")) for k in ("fwd", "bwd"))


def _lib(kind: str):
    if kind not in _LIBS:
        path = os.path.join(OUT, f"libref_trace_{kind}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `python oracle/build_ref.py` in the build container")
        L = ctypes.CDLL(path)
        L.ref_trace_launch.argtypes = [ctypes.POINTER(_Args), ctypes.c_void_p]
        L.ref_trace_launch.restype = ctypes.c_int
        assert L.ref_trace_is_backward() == (1 if kind.startswith("bwd") else 0)
        _LIBS[kind] = L
    return _LIBS[kind]


# ---- real OptiX backend (oracle/ref_trace_optix_driver.cpp) -----------------------------------------------------------
def _optix_init():
    if _OPTIX:
        return
    try:
        import glob
        so = os.path.join(OPTIX_OUT, "libref_trace_optix.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        # libnvoptix and the RTX core it dlopens by name sit outside the default loader path on this pool's boxes
        # (/usr/local/nvidia/lib): preload both by absolute path so the later dlopen-by-soname calls resolve
        for dep in sorted(glob.glob("/usr/local/nvidia/lib*/libnvidia-rtcore.so.*")) + sorted(glob.glob("/usr/local/nvidia/lib*/libnvidia-ptxjitcompiler.so.*")):
            try:
                ctypes.CDLL(dep, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                pass
        loaded = None
        for c in ["libnvoptix.so.1"] + sorted(glob.glob("/usr/local/nvidia/lib*/libnvoptix.so.1")) + \
                sorted(glob.glob("/usr/lib/x86_64-linux-gnu/libnvoptix.so.1")):
            try:
                ctypes.CDLL(c, mode=ctypes.RTLD_GLOBAL)
                loaded = c
                break
            except OSError:
                continue
        if loaded is None:
            raise OSError("libnvoptix.so.1 not loadable")
        L = ctypes.CDLL(so)
        L.ref_optix_last_error.restype = ctypes.c_char_p
        L.ref_optix_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.ref_optix_build.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ref_optix_launch.argtypes = [ctypes.POINTER(_Args), ctypes.c_int, ctypes.c_void_p]
        last = ""
        for arch in ("compute_100", "compute_80"):
            d = os.path.join(OPTIX_OUT, "ptx_" + arch)
            if L.ref_optix_create(os.path.join(d, "forward.ptx").encode(), os.path.join(d, "backward.ptx").encode(), 0) == 0:
                _OPTIX.update(lib=L, ptx=arch, nvoptix=loaded)
                return
            last = L.ref_optix_last_error().decode()
        raise RuntimeError(f"OptiX pipeline creation failed: {last}")
    except Exception as ex:   # noqa: BLE001
        _OPTIX["err"] = f"{type(ex).__name__}: {ex}"


def optix_available() -> bool:
    if not torch.cuda.is_available():
        return False
    _optix_init()
    return "lib" in _OPTIX


def optix_info() -> dict:
    _optix_init()
    return {k: _OPTIX.get(k) for k in ("ptx", "nvoptix", "err")}


def optix_build(vertices: torch.Tensor, triangles: torch.Tensor) -> None:
    """optixAccelBuild of the quad triangles (trace_surfels.cpp:47-121); the handle lives in the driver."""
    _optix_init()
    L = _OPTIX["lib"]
    v, t = vertices.detach().contiguous().float(), triangles.contiguous().int()
    with torch.cuda.device(v.device):
        rc = L.ref_optix_build(v.data_ptr(), v.shape[0], t.data_ptr(), t.shape[0], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.current_stream().synchronize()
    if rc != 0:
        raise RuntimeError("reference OptiX build failed: " + L.ref_optix_last_error().decode())
    _OPTIX["built_for"] = (v.data_ptr(), v.shape[0])


def _launch(backend: str, backward: bool, a, kw):
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if backend == "optix":
        _optix_init()
        if "lib" not in _OPTIX:
            raise RuntimeError("real-OptiX reference backend unavailable: " + _OPTIX.get("err", ""))
        if kw.get("rebuild", True):
            optix_build(kw["vertices"], kw["triangles"])
        L = _OPTIX["lib"]
        rc = L.ref_optix_launch(ctypes.byref(a), int(backward), stream)
        if rc != 0:
            raise RuntimeError("reference OptiX launch failed: " + L.ref_optix_last_error().decode())
        return
    kind = ("bwd" if backward else "fwd") + ("_nudge" if backend == "emulation_nudge" else "")
    rc = _lib(kind).ref_trace_launch(ctypes.byref(a), stream)
    if rc != 0:
        raise RuntimeError(f"reference tracer {kind} failed: {rc}")


def _ptr(t):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


def _common(a, keep, *, ray_o, ray_d, vertices, triangles, bg, means3D, shs, degree, colors_precomp, others_precomp, opacities,
            scales, scale_modifier, rotations, viewmatrix, projmatrix, campos, start_from_first, max_trace_depth,
            specular_threshold, rebuild=True):
    f32 = lambda t: None if t is None else t.detach().contiguous().float()
    ray_o, ray_d, vertices, bg, means3D, shs, colors_precomp, others_precomp, opacities, scales, rotations, viewmatrix, \
        projmatrix, campos = map(f32, (ray_o, ray_d, vertices, bg, means3D, shs, colors_precomp, others_precomp, opacities,
                                       scales, rotations, viewmatrix, projmatrix, campos))
    triangles = triangles.contiguous().int()
    keep += [ray_o, ray_d, vertices, triangles, bg, means3D, shs, colors_precomp, others_precomp, opacities, scales, rotations,
             viewmatrix, projmatrix, campos]
    a.P, a.H, a.W, a.D = means3D.shape[0], ray_o.shape[0], ray_o.shape[1], int(degree)
    a.M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
    a.start_from_first, a.max_trace_depth = int(bool(start_from_first)), int(max_trace_depth)
    a.specular_threshold, a.scale_modifier = float(specular_threshold), float(scale_modifier)
    a.num_triangles = triangles.shape[0]
    for name, t in dict(vertices=vertices, triangles=triangles, ray_o=ray_o, ray_d=ray_d, background=bg, means3D=means3D, shs=shs,
                        colors_precomp=colors_precomp, others_precomp=others_precomp, opacities=opacities, scales=scales,
                        rotations=rotations, viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos).items():
        setattr(a, name, _ptr(t))
    return a.P, a.H, a.W, a.M


def trace_forward(*, training=True, backend="emulation", **kw):
    """-> dict(rgb, dpt, acc, norm, dist(H,W,3), aux, mid, wet) of CUDA tensors.  backend: "emulation" (brute-force OptiX
    emulation), "emulation_nudge" (same, hit distances moved by one ulp), "optix" (the driver's real OptiX runtime)."""
    a, keep = _Args(), []
    P, H, W, M = _common(a, keep, **kw)
    dev = kw["means3D"].device
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    out = dict(rgb=z(H, W, 3), dpt=z(H, W, 1), acc=z(H, W, 1), norm=z(H, W, 3), dist=z(H, W, 3), aux=z(H, W, 2), mid=z(H, W, 32),
               wet=z(P, 1))
    a.training = int(bool(training))
    a.out_rgb, a.out_dpt, a.out_acc, a.out_norm, a.out_dist, a.out_aux, a.mid_val = (
        _ptr(out[k]) for k in ("rgb", "dpt", "acc", "norm", "dist", "aux", "mid"))
    a.a_weights = _ptr(out["wet"]) if training else None
    with torch.cuda.device(dev):
        _launch(backend, False, a, dict(kw, vertices=keep[2], triangles=keep[3]))
    return out


def trace_backward(fwd: dict, grads: dict, backend="emulation", **kw):
    """fwd = trace_forward's result; grads = dict(rgb, dpt, acc, norm, dist(H,W,1), aux) upstream gradients.
    -> dict(g_ray_o, g_ray_d, g_means, g_shs, g_colors, g_others, g_opac, g_scales, g_rots, g_transMat)."""
    a, keep = _Args(), []
    P, H, W, M = _common(a, keep, **kw)
    dev = kw["means3D"].device
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    g = {k: v.detach().to(dev).contiguous().float() for k, v in grads.items()}
    out = dict(g_ray_o=z(H, W, 3), g_ray_d=z(H, W, 3), g_means=z(P, 3), g_shs=z(P, M, 3), g_colors=z(P, 3), g_others=z(P, 2),
               g_opac=z(P, 1), g_scales=z(P, 2), g_rots=z(P, 4), g_transMat=z(P, 9))
    a.out_rgb, a.out_dpt, a.out_acc, a.out_norm, a.out_dist, a.out_aux, a.mid_val = (
        _ptr(fwd[k]) for k in ("rgb", "dpt", "acc", "norm", "dist", "aux", "mid"))
    a.dL_drgb, a.dL_ddpt, a.dL_dacc, a.dL_ddist, a.dL_dnorm, a.dL_daux = (
        _ptr(g[k]) for k in ("rgb", "dpt", "acc", "dist", "norm", "aux"))
    a.dL_dray_o, a.dL_dray_d, a.dL_dmeans3D, a.dL_dshs, a.dL_dcolors, a.dL_dothers, a.dL_dopacities, a.dL_dscales, \
        a.dL_drotations, a.dL_dtransMat = (ctypes.c_void_p(out[k].data_ptr()) for k in (
            "g_ray_o", "g_ray_d", "g_means", "g_shs", "g_colors", "g_others", "g_opac", "g_scales", "g_rots", "g_transMat"))
    with torch.cuda.device(dev):
        _launch(backend, True, a, dict(kw, vertices=keep[2], triangles=keep[3], rebuild=False))
    return out
