"""ORACLE tooling (checker / timing baseline only; never imported by the product path): the UNMODIFIED reference operators
built under oracle/_ref, wrapped in autograd Functions that call them the way the reference's own Python wrappers do
    diff_surfel_rasterization_wet_ch05/__init__.py:44-156   (_RasterizeGaussians)
    diff_surfel_tracing/__init__.py:51-204                  (_TraceSurfels)
so that `bench.py --impl reference` and bench.py's `ref_gpu` block can time the reference's rasterizer (its CUDA extension
compiled for sm_100a) and tracer (its device programs under the driver's real OptiX runtime) on the same B200, same inputs,
same process.  `available()` says whether both can run here."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import build_ref
from . import ref_trace as RT

_MODS: dict = {}


def _rast_mod(name="ref_rast_wet_ch05"):
    if name not in _MODS:
        _MODS[name] = build_ref.load(name)
    return _MODS[name]


def available() -> tuple[bool, str]:
    try:
        _rast_mod()
    except Exception as ex:   # noqa: BLE001
        return False, f"reference rasterizer extension not loadable: {ex}"
    if not RT.optix_available():
        return False, "reference tracer needs the real OptiX runtime: " + str(RT.optix_info().get("err"))
    return True, ""


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors, opacities, scales, rotations, s):
        _C = _rast_mod()
        e = torch.empty(0, device=means3D.device)
        n, color, allmap, weight, radii, gb, bb, ib = _C.rasterize_gaussians(
            s.bg, means3D, colors, opacities, scales, rotations, s.scale_modifier, e, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy,
            s.image_height, s.image_width, e, s.sh_degree, s.campos, s.prefiltered, s.debug)
        ctx.s, ctx.n = s, n
        ctx.save_for_backward(colors, means3D, scales, rotations, radii, gb, bb, ib)
        ctx.mark_non_differentiable(radii)
        return color, radii, allmap, weight

    @staticmethod
    def backward(ctx, g_color, g_radii, g_allmap, g_weight):
        _C, s = _rast_mod(), ctx.s
        colors, means3D, scales, rotations, radii, gb, bb, ib = ctx.saved_tensors
        e = torch.empty(0, device=means3D.device)
        g = _C.rasterize_gaussians_backward(s.bg, means3D, radii, colors, scales, rotations, s.scale_modifier, e, s.viewmatrix, s.projmatrix,
                                            s.tanfovx, s.tanfovy, g_color.contiguous(), g_allmap.contiguous(), e, s.sh_degree, s.campos, gb,
                                            ctx.n, bb, ib, s.debug)
        g_means2D, g_colors, g_opac, g_means3D, g_cov, g_sh, g_scales, g_rots = g
        return g_means3D, g_means2D, g_colors, g_opac, g_scales, g_rots, None


class RefGaussianRasterizer(nn.Module):
    """Drop-in for envgs_b200.rasterizer's GaussianRasterizer (C = 5, precomputed colours) backed by the reference build.
    The reference reads the background out of bounds when handed 3 floats for 5 channels; pad it like the product shim does."""

    def __init__(self, raster_settings):
        super().__init__()
        bg = raster_settings.bg
        if bg.numel() < 5:
            bg = torch.cat([bg.reshape(-1), bg.new_zeros(5 - bg.numel())])
        self.s = raster_settings._replace(bg=bg.contiguous())

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        return _RefRasterize.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, self.s)


class _RefTrace(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ray_o, ray_d, vertices, triangles, means3D, grads3D, shs, opacities, scales, rotations, s, start_from_first):
        kw = dict(ray_o=ray_o, ray_d=ray_d, vertices=vertices, triangles=triangles, bg=s.bg, means3D=means3D, shs=shs, degree=s.sh_degree,
                  colors_precomp=None, others_precomp=None, opacities=opacities, scales=scales, scale_modifier=s.scale_modifier,
                  rotations=rotations, viewmatrix=s.viewmatrix, projmatrix=s.projmatrix, campos=s.campos,
                  start_from_first=start_from_first, max_trace_depth=s.max_trace_depth, specular_threshold=s.specular_threshold)
        fwd = RT.trace_forward(training=True, backend="optix", rebuild=False, **kw)
        ctx.kw, ctx.fwd = kw, fwd
        return fwd["rgb"], fwd["dpt"], fwd["acc"], fwd["norm"], fwd["dist"][..., :1].clone(), fwd["aux"], fwd["mid"], fwd["wet"]

    @staticmethod
    def backward(ctx, g_rgb, g_dpt, g_acc, g_norm, g_dist, g_aux, g_mid, g_wet):
        kw = ctx.kw
        z = lambda t, like: torch.zeros_like(like) if t is None else t
        f = ctx.fwd
        up = dict(rgb=z(g_rgb, f["rgb"]), dpt=z(g_dpt, f["dpt"]), acc=z(g_acc, f["acc"]), norm=z(g_norm, f["norm"]),
                  dist=z(g_dist, f["dpt"]), aux=z(g_aux, f["aux"]))
        bw = RT.trace_backward(f, up, backend="optix", **kw)
        scale = torch.norm(kw["means3D"] - kw["campos"][None], dim=-1, keepdim=True)
        return (bw["g_ray_o"], bw["g_ray_d"], None, None, bw["g_means"], bw["g_means"] * scale / 2, bw["g_shs"], bw["g_opac"], bw["g_scales"],
                bw["g_rots"], None, None)


class RefSurfelTracer(nn.Module):
    """Drop-in for envgs_b200.tracer.SurfelTracer backed by the reference device programs under real OptiX."""

    def build_acceleration_structure(self, vertices, triangles, rebuild=1):
        self._tri = triangles
        RT.optix_build(vertices, triangles)

    def forward(self, ray_o, ray_d, vertices, means3D, grads3D, shs=None, colors_precomp=None, others_precomp=None, opacities=None,
                scales=None, rotations=None, cov3D_precomp=None, tracer_settings=None, start_from_first=True):
        return _RefTrace.apply(ray_o, ray_d, vertices, self._tri, means3D, grads3D, shs, opacities, scales, rotations, tracer_settings,
                               start_from_first)
