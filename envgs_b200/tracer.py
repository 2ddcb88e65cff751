"""Python operator surface of the surfel tracer, mirroring the reference package `diff_surfel_tracing`
(TRACE/diff_surfel_tracing/__init__.py:13-293): `SurfelTracingSettings`, `SurfelTracer` with
`build_acceleration_structure(vertices, triangles, rebuild)` and `forward(ray_o, ray_d, vertices, means3D, grads3D, ...)`
returning `(rgb, dpt, acc, norm, dist[..., :1], aux, mid, wet)`.  OptiX is replaced by the LBVH tracer in
libenvgs_b200.so (`envgs_tracer_*`, `envgs_trace_forward/backward`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from . import gradsink

NUM_CHANNELS = 3
AUX_CHANNELS = 2
MID_CHANNELS = 16
MAX_TRACE_DEPTH = 1


class SurfelTracingSettings(NamedTuple):
    image_height: int  # no use, only for compatibility
    image_width: int  # no use, only for compatibility
    tanfovx: float  # no use, only for compatibility
    tanfovy: float  # no use, only for compatibility
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    max_trace_depth: int
    specular_threshold: float


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


class _Context:
    """Owns the native tracer handle and the torch-allocated BVH storage (the OptiXStateWrapper stand-in)."""

    def __init__(self):
        self._lib = _lib.lib()
        self._handle = None          # created lazily, under the device guard of the tensors it will serve
        self.storage = None
        self.num_surfels = 0
        self._cb = _lib.ALLOC_FN(self._alloc)
        self._device = None

    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("the acceleration structure has not been built: call build_acceleration_structure first")
        return self._handle

    def _ensure(self, device: torch.device):
        """The native handle owns a work counter, a side stream and two events: they belong to ONE CUDA device, so it is
        created under `torch.cuda.device(device)` and re-created if the module is later used on another device."""
        if self._handle is not None and self._device == device:
            return
        self._destroy()
        with torch.cuda.device(device):
            self._handle = self._lib.envgs_tracer_create()
        if not self._handle:
            raise RuntimeError("envgs_tracer_create failed: " + self._lib.envgs_last_error().decode())
        self._device = device
        self.storage, self.num_surfels = None, 0

    def _destroy(self):
        if self._handle:
            if self._device is not None:
                with torch.cuda.device(self._device):
                    self._lib.envgs_tracer_destroy(self._handle)
            else:
                self._lib.envgs_tracer_destroy(self._handle)
            self._handle = None

    def _alloc(self, _user, nbytes):
        try:
            self.storage = torch.empty(int(nbytes), dtype=torch.uint8, device=self._device)
        except Exception:
            return None
        return self.storage.data_ptr()

    def build(self, vertices: torch.Tensor, rebuild: int):
        if vertices.dim() != 2 or vertices.shape[1] != 3:
            raise RuntimeError("vertices must have dimensions (num_vertices, 3)")
        if not vertices.is_cuda:
            raise RuntimeError("vertices must be a CUDA tensor")
        v = vertices.detach().to(torch.float32).contiguous()
        self._ensure(v.device)
        if v.shape[0] == 0:          # a set pruned to zero surfels: nothing to build, forward returns the background
            self.num_surfels = 0
            return
        with torch.cuda.device(v.device):
            _lib.check(self._lib.envgs_tracer_build(self.handle, _lib.ptr(v), v.shape[0], int(rebuild), self._cb, None,
                                                    torch.cuda.current_stream(v.device).cuda_stream),
                       "envgs_tracer_build")
        self.num_surfels = v.shape[0] // 4

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


def _f32c(t, device):
    if t is None:
        return None
    if t.device != device:
        raise RuntimeError("all tracer inputs must be CUDA tensors on the same device")
    return t.detach().to(torch.float32).contiguous()


def _trace_args(P, H, W, s, training, start_from_first, ray_o, ray_d, bg, means3D, shs, colors, others, opacities, scales,
                rotations, vm, pm, cam):
    a = _lib.TraceArgs()
    a.P, a.H, a.W = P, H, W
    a.sh_degree = int(s.sh_degree)
    a.sh_coeffs = shs.shape[1] if shs.numel() else 0
    a.training = int(bool(training))
    a.start_from_first = int(bool(start_from_first))
    a.max_trace_depth = int(s.max_trace_depth)
    a.specular_threshold = float(s.specular_threshold)
    a.scale_modifier = float(s.scale_modifier)
    a.ray_o, a.ray_d, a.background = _lib.ptr(ray_o), _lib.ptr(ray_d), _lib.ptr(bg)
    a.means3D, a.shs, a.colors_precomp = _lib.ptr(means3D), _lib.ptr(shs), _lib.ptr(colors)
    a.others_precomp, a.opacities = _lib.ptr(others), _lib.ptr(opacities)
    a.scales, a.rotations = _lib.ptr(scales), _lib.ptr(rotations)
    a.viewmatrix, a.projmatrix, a.campos = _lib.ptr(vm), _lib.ptr(pm), _lib.ptr(cam)
    return a


class _TraceSurfels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, optix_context, training, start_from_first, ray_o, ray_d, vertices, means3D, grads3D, shs,
                colors_precomp, others_precomp, opacities, scales, rotations, cov3Ds_precomp, tracer_settings):
        L = _lib.lib()
        s = tracer_settings
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if cov3Ds_precomp is not None and cov3Ds_precomp.numel():
            raise RuntimeError("precomputed 3D covariance is not supported by the tracer (the reference ignores it too)")
        if ray_o.dim() != 3 or ray_o.shape[-1] != 3 or ray_o.shape != ray_d.shape:
            raise RuntimeError("ray_o / ray_d must both have dimensions (H, W, 3)")
        device = means3D.device
        P = means3D.shape[0]
        H, W = ray_o.shape[0], ray_o.shape[1]
        if optix_context._device is not None and optix_context._device != device:
            raise RuntimeError(f"the acceleration structure was built on {optix_context._device} but the surfels are on {device}")
        if P == 0:
            # an empty traced set (e.g. pruned to zero surfels): every ray sees the background; no gradient flows anywhere
            f = dict(dtype=torch.float32, device=device)
            bg = s.bg.detach().to(**f).reshape(-1)[:3]
            rgb = bg.expand(H, W, 3).contiguous()
            mid = torch.zeros((H, W, MID_CHANNELS * (MAX_TRACE_DEPTH + 1)), **f)
            mid[..., 0:3], mid[..., 3:6], mid[..., 13:16] = ray_o.detach(), ray_d.detach(), rgb
            ctx.empty = True
            wet = torch.zeros((0, 1), **f)
            ctx.mark_non_differentiable(mid, wet)
            return (rgb, torch.zeros((H, W, 1), **f), torch.zeros((H, W, 1), **f), torch.zeros((H, W, 3), **f), torch.zeros((H, W, 1), **f),
                    torch.zeros((H, W, AUX_CHANNELS), **f), mid, wet)
        ctx.empty = False

        ro, rd = _f32c(ray_o, device), _f32c(ray_d, device)
        m3, sh, col, oth = _f32c(means3D, device), _f32c(shs, device), _f32c(colors_precomp, device), _f32c(others_precomp, device)
        opa, sc, rot = _f32c(opacities, device), _f32c(scales, device), _f32c(rotations, device)
        bg = s.bg.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        vm, pm, cam = _f32c(s.viewmatrix, device), _f32c(s.projmatrix, device), _f32c(s.campos, device)

        f = dict(dtype=torch.float32, device=device)
        # every ray's outputs are written by the kernels
        out_rgb, out_dpt, out_acc = torch.empty((H, W, 3), **f), torch.empty((H, W, 1), **f), torch.empty((H, W, 1), **f)
        out_norm, out_dist, out_aux = torch.empty((H, W, 3), **f), torch.empty((H, W, 3), **f), torch.empty((H, W, AUX_CHANNELS), **f)
        mid_val = torch.empty((H, W, MID_CHANNELS * (MAX_TRACE_DEPTH + 1)), **f)   # fully written by the kernel
        a_weights = torch.zeros((P, 1), **f) if training else None
        # scratch for the hit-parallel fast path (per-ray hit lists, re-used by backward); the library falls back to the
        # general kernel when the call shape is not eligible
        hit_cache = None
        depth = int(s.max_trace_depth)
        if not start_from_first and (depth == 0 or (depth == 1 and oth is not None and oth.numel())) \
                and os.environ.get("ENVGS_TRACE_NO_CACHE") != "1":
            hit_cache = torch.empty(int(L.envgs_trace_cache_bytes2(H, W, depth)), dtype=torch.uint8, device=device)

        args = _trace_args(P, H, W, s, training, start_from_first, ro, rd, bg, m3, sh, col, oth, opa, sc, rot, vm, pm, cam)
        out = _lib.TraceOut(_lib.ptr(out_rgb), _lib.ptr(out_dpt), _lib.ptr(out_acc), _lib.ptr(out_norm), _lib.ptr(out_dist),
                            _lib.ptr(out_aux), _lib.ptr(mid_val), _lib.ptr(a_weights), _lib.ptr(hit_cache))
        stream = torch.cuda.current_stream(device).cuda_stream

        def run():
            with torch.cuda.device(device):
                _lib.check(L.envgs_trace_forward(optix_context.handle, C.byref(args), C.byref(out), stream),
                           "envgs_trace_forward")

        if s.debug:
            cpu_args = cpu_deep_copy_tuple((training, start_from_first, ray_o, ray_d, vertices, s.bg, means3D, shs, s.sh_degree,
                                            colors_precomp, others_precomp, opacities, scales, s.scale_modifier, rotations,
                                            cov3Ds_precomp, s.viewmatrix, s.projmatrix, s.campos, s.prefiltered, s.debug,
                                            s.max_trace_depth, s.specular_threshold))
            try:
                run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            run()

        ctx.tracer_settings = s
        ctx.optix_context = optix_context
        ctx.start_from_first = start_from_first
        ctx.bg = bg
        ctx.hit_cache = hit_cache
        ctx.save_for_backward(ro, rd, m3, sh, col, oth, opa, sc, rot, out_rgb, out_dpt, out_acc, out_norm, out_dist,
                              out_aux, mid_val)
        acc_wet = a_weights if a_weights is not None else torch.zeros_like(means3D[:, 0])
        ctx.mark_non_differentiable(mid_val, acc_wet)
        return out_rgb, out_dpt, out_acc, out_norm, out_dist[..., :1].clone(), out_aux, mid_val, acc_wet

    @staticmethod
    def backward(ctx, grad_out_rgb, grad_out_dpt, grad_out_acc, grad_out_norm, grad_out_dist, grad_out_aux, grad_mid_val,
                 grad_acc_wet):
        if ctx.empty:               # nothing was traced: every input's gradient is zero (None)
            return (None,) * 16
        L = _lib.lib()
        s = ctx.tracer_settings
        ro, rd, m3, sh, col, oth, opa, sc, rot, out_rgb, out_dpt, out_acc, out_norm, out_dist, out_aux, mid_val = ctx.saved_tensors
        device = m3.device
        P = m3.shape[0]
        H, W = ro.shape[0], ro.shape[1]
        M = sh.shape[1] if sh.numel() else 0
        f = dict(dtype=torch.float32, device=device)

        def up(g, shape):
            return torch.zeros(shape, **f) if g is None else g.to(torch.float32).contiguous()

        g_rgb, g_dpt, g_acc = up(grad_out_rgb, (H, W, 3)), up(grad_out_dpt, (H, W, 1)), up(grad_out_acc, (H, W, 1))
        g_norm, g_dist, g_aux = up(grad_out_norm, (H, W, 3)), up(grad_out_dist, (H, W, 1)), up(grad_out_aux, (H, W, AUX_CHANNELS))

        dL_dray_o, dL_dray_d = torch.empty((H, W, 3), **f), torch.empty((H, W, 3), **f)
        zs = lambda saved, *shape: gradsink.zeros_or_sink(saved, shape, **f)     # the trainer's flat-buffer slice if registered
        dL_dmeans3D = zs(m3, P, 3)
        dL_dshs = zs(sh, P, M, 3) if M else torch.zeros((P, M, 3), **f)
        dL_dcolors = torch.zeros((P, NUM_CHANNELS), **f)
        dL_dothers, dL_dopacities = torch.zeros((P, AUX_CHANNELS), **f), zs(opa, P, 1)
        dL_dscales, dL_drotations, dL_dtransMat = zs(sc, P, 2), zs(rot, P, 4), torch.zeros((P, 9), **f)

        vm, pm, cam = _f32c(s.viewmatrix, device), _f32c(s.projmatrix, device), _f32c(s.campos, device)
        args = _trace_args(P, H, W, s, False, ctx.start_from_first, ro, rd, ctx.bg, m3, sh, col, oth, opa, sc, rot, vm, pm, cam)
        grads = _lib.TraceGrads(
            _lib.ptr(out_rgb), _lib.ptr(out_dpt), _lib.ptr(out_acc), _lib.ptr(out_norm), _lib.ptr(out_dist), _lib.ptr(out_aux),
            _lib.ptr(mid_val), _lib.ptr(g_rgb), _lib.ptr(g_dpt), _lib.ptr(g_acc), _lib.ptr(g_norm), _lib.ptr(g_dist),
            _lib.ptr(g_aux), _lib.ptr(dL_dray_o), _lib.ptr(dL_dray_d), _lib.ptr(dL_dmeans3D), _lib.ptr(dL_dshs),
            _lib.ptr(dL_dcolors), _lib.ptr(dL_dothers), _lib.ptr(dL_dopacities), _lib.ptr(dL_dscales), _lib.ptr(dL_drotations),
            _lib.ptr(ctx.hit_cache))
        stream = torch.cuda.current_stream(device).cuda_stream

        def run():
            with torch.cuda.device(device):
                _lib.check(L.envgs_trace_backward(ctx.optix_context.handle, C.byref(args), C.byref(grads), stream),
                           "envgs_trace_backward")

        if s.debug:
            cpu_args = cpu_deep_copy_tuple((ctx.start_from_first, ro, rd, s.bg, m3, sh, s.sh_degree, col, oth, opa, sc,
                                            s.scale_modifier, rot, s.viewmatrix, s.projmatrix, s.campos, s.max_trace_depth,
                                            s.specular_threshold, out_rgb, out_dpt, out_acc, out_norm, out_dist, out_aux,
                                            mid_val, grad_out_rgb, grad_out_dpt, grad_out_acc, grad_out_norm, grad_out_dist,
                                            grad_out_aux))
            try:
                run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            run()

        # densification signal for the traced set (TRACE/diff_surfel_tracing/__init__.py:182-183)
        grad_scale = torch.norm(m3 - cam[None], dim=-1, keepdim=True)
        grad_grads3D = dL_dmeans3D * grad_scale / 2

        like = lambda g, ref: g if ref.numel() else None
        return (None, None, None, dL_dray_o, dL_dray_d, None, dL_dmeans3D, grad_grads3D, like(dL_dshs, sh),
                like(dL_dcolors, col), like(dL_dothers, oth), dL_dopacities, dL_dscales, dL_drotations, None, None)


def trace_surfels(optix_context, training, start_from_first, ray_o, ray_d, vertices, means3D, grads3D, shs, colors_precomp,
                  others_precomp, opacities, scales, rotations, cov3Ds_precomp, tracer_settings):
    return _TraceSurfels.apply(optix_context, training, start_from_first, ray_o, ray_d, vertices, means3D, grads3D, shs,
                               colors_precomp, others_precomp, opacities, scales, rotations, cov3Ds_precomp, tracer_settings)


class SurfelTracer(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        # The reference creates an OptiX context + pipelines here (optix_wrapper.cpp:177-215); this one owns an LBVH.
        self.optix_context = _Context()

    def build_acceleration_structure(self, vertices: torch.Tensor, triangles: torch.Tensor, rebuild: bool = 1):
        if triangles is not None:
            if triangles.dim() != 2 or triangles.shape[1] != 3:
                raise RuntimeError("triangles must have dimensions (num_triangles, 3)")
            if triangles.shape[0] * 2 != vertices.shape[0]:
                raise RuntimeError("expected two triangles per four vertices (one quad per surfel, optix_utils.py:39-69)")
        return self.optix_context.build(vertices, rebuild)

    def forward(self, ray_o, ray_d, vertices, means3D, grads3D, shs=None, colors_precomp=None, others_precomp=None,
                opacities=None, scales=None, rotations=None, cov3D_precomp=None, tracer_settings=None,
                start_from_first: bool = True):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = lambda: torch.empty(0, dtype=torch.float32, device=means3D.device)
        if shs is None: shs = empty()
        if colors_precomp is None: colors_precomp = empty()
        if others_precomp is None: others_precomp = empty()
        if scales is None: scales = empty()
        if rotations is None: rotations = empty()
        if cov3D_precomp is None: cov3D_precomp = empty()
        if vertices is None: vertices = empty()
        return trace_surfels(self.optix_context, self.training, start_from_first, ray_o, ray_d, vertices, means3D, grads3D,
                             shs, colors_precomp, others_precomp, opacities, scales, rotations, cov3D_precomp,
                             tracer_settings)
