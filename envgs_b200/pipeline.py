"""Per-view EnvGS rendering path built on the two operators — the host-side mirror of
`EnvGSSampler.forward` (easyvolcap/models/samplers/envgs_sampler.py:482-565), `render()`
(easyvolcap/utils/gaussian2d_utils.py:1003-1155) and `HardwareRendering.render_gaussians`
(easyvolcap/utils/optix_utils.py:87-267), reduced to what the hot path needs:

    base surfels --rasterize (C=5)--> rgb/spec/rough/depth/normal --reflect--> rays --trace env surfels--> blend

This is the public API `bench.py` times end to end and what the view-parallel driver (`envgs_b200.dist`) replicates per
rank.  All map-sized glue stays PyTorch, exactly as in the reference (the drop-in boundary is below it).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import scene
from . import glue
from .features import sh_features
from .rasterizer import GaussianRasterizationSettings, make_module
from .tracer import SurfelTracer, SurfelTracingSettings

_, _, _Rasterizer5 = make_module(5)

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def eval_sh_deg3(sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """(P,16,3) coefficients, (P,3) unit directions -> (P,3); same basis as easyvolcap/utils/sh_utils.py:eval_sh."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    res = SH_C0 * sh[:, 0] - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    res = res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6] \
        + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8]
    res = res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] \
        + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] \
        + SH_C3[5] * z * (xx - yy) * sh[:, 14] + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return res


def depth_to_normal(cam: scene.Camera, dpt: torch.Tensor) -> torch.Tensor:
    """(1,H,W) depth -> (H,W,3) pseudo surface normal (gaussian2d_utils.py:1158-1206)."""
    H, W = cam.H, cam.W
    dev = dpt.device
    c2w = torch.linalg.inv(cam.viewmatrix.t())
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    Kinv = torch.tensor([[1 / fx, 0, -(W / 2) / fx], [0, 1 / fy, -(H / 2) / fy], [0, 0, 1.0]], device=dev)
    u, v = torch.meshgrid(torch.arange(W, device=dev, dtype=torch.float32), torch.arange(H, device=dev, dtype=torch.float32),
                          indexing="xy")
    xyz = torch.stack([u, v, torch.ones_like(u)], dim=-1).reshape(-1, 3)
    ray_d = xyz @ Kinv.t() @ c2w[:3, :3].t()
    pts = (dpt.reshape(-1, 1) * ray_d + c2w[:3, 3]).reshape(H, W, 3)
    out = torch.zeros_like(pts)
    dx = pts[2:, 1:-1] - pts[:-2, 1:-1]
    dy = pts[1:-1, 2:] - pts[1:-1, :-2]
    out[1:-1, 1:-1] = torch.nn.functional.normalize(torch.linalg.cross(dx, dy, dim=-1), dim=-1)
    return out


class EnvGSRenderer(nn.Module):
    """Holds the base and environment surfel sets as parameters and renders / differentiates one view."""

    def __init__(self, base: scene.SurfelSet, env: scene.SurfelSet | None, device="cuda", fused_features: bool = True,
                 fused_glue: bool = True, rasterizer_cls=None, tracer=None, max_trace_depth: int = 0, specular_threshold: float = 0.01,
                 ray_filter: str | None = None, specular_filtering_percent: float = 0.75):
        """`rasterizer_cls` / `tracer`: operator objects with the reference's Python surface (GaussianRasterizer class for 5
        channels, SurfelTracer instance).  Default: this package's.  bench.py's reference arm passes the wrappers of the
        reference build (oracle/ref_ops.py) together with fused_* = False, i.e. the reference's own PyTorch glue."""
        super().__init__()
        self._rasterizer_cls = rasterizer_cls if rasterizer_cls is not None else _Rasterizer5
        # max_trace_depth = 1: the traced set carries its own specular / roughness (`others_precomp`) and a ray continues along
        # the reflection about the accumulated normal while the accumulated specular exceeds the threshold (BASELINE config 5;
        # TRACE/optix_tracer/forward.cu:453-559).  EnvGS' own configuration traces a single segment (0).
        self.max_trace_depth, self.specular_threshold = int(max_trace_depth), float(specular_threshold)
        # Which pixels send a reflection ray (envgs_sampler.py:433-447): None = all (the reference's default configuration),
        # "acc" = accumulated alpha > 0.75, "specular" = rendered specular above its `specular_filtering_percent` quantile — the
        # reference's two optional filters, traced as a compact (1, S) ray list exactly like there — and "nonempty" = pixels with
        # alpha > 0: a pixel no base surfel reaches has specular == 0 exactly, so its environment colour is multiplied by zero
        # and no parameter gradient can depend on it; image, loss and every parameter gradient are unchanged, only the traced
        # set's blend-weight statistics (`wet`) lose those rays' contributions.
        assert ray_filter in (None, "acc", "specular", "nonempty")
        self.ray_filter, self.specular_filtering_percent = ray_filter, specular_filtering_percent
        self.fused_features = fused_features   # False: evaluate the base SH colours in PyTorch exactly like the reference
        self.fused_glue = fused_glue           # False: per-pixel glue (reflection, dpt2norm, blend, loss) in PyTorch like the reference
        self.lambda_normal = 0.04
        # fused path: build / refit the environment set's acceleration structure on a side stream underneath the rasterizer forward
        self.async_bvh = os.environ.get("ENVGS_ASYNC_BVH", "1") != "0"
        self._env_stream = None
        dev = torch.device(device)
        P = lambda t: nn.Parameter(t.to(dev).contiguous())
        self.means, self.scales, self.rotations, self.opacities = P(base.means), P(base.scales), P(base.rotations), P(base.opacities)
        self.shs, self.specular, self.roughness = P(base.shs), P(base.specular), P(base.roughness)
        self.has_env = env is not None
        if self.has_env:
            self.env_means, self.env_scales, self.env_rotations = P(env.means), P(env.scales), P(env.rotations)
            self.env_opacities, self.env_shs = P(env.opacities), P(env.shs)
            if self.max_trace_depth > 0:
                self.env_specular, self.env_roughness = P(env.specular), P(env.roughness)
            self.tracer = tracer if tracer is not None else SurfelTracer()
        self.register_buffer("bg", torch.zeros(3, device=dev))
        self.register_buffer("env_bg", torch.zeros(3, device=dev))

    # ---- base pass -------------------------------------------------------------------------------------------
    def _raster(self, cam: scene.Camera):
        settings = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, self.bg, 1.0, cam.viewmatrix,
                                                 cam.projmatrix, 3, cam.campos, False, False)
        if self.fused_features:
            colors = sh_features(self.means, self.shs, torch.cat([self.specular, self.roughness], dim=-1), cam.campos, 3)
        else:
            dirs = torch.nn.functional.normalize(self.means - cam.campos[None], dim=-1)
            rgb = torch.clamp_min(eval_sh_deg3(self.shs, dirs) + 0.5, 0.0)
            colors = torch.cat([rgb, self.specular, self.roughness], dim=-1)
        means2D = torch.zeros_like(self.means, requires_grad=True)
        image, radii, allmap, weight = self._rasterizer_cls(settings)(means3D=self.means, means2D=means2D, opacities=self.opacities,
                                                              colors_precomp=colors, scales=self.scales, rotations=self.rotations)
        return image, radii, allmap, weight, means2D

    def rasterize_base(self, cam: scene.Camera) -> dict:
        image, radii, allmap, weight, means2D = self._raster(cam)
        alpha = allmap[1:2]
        Rv = cam.viewmatrix[:3, :3]                                                              # view -> world: n_w = Rv n_v
        nv = allmap[2:5]
        normal = torch.stack([Rv[0, 0] * nv[0] + Rv[0, 1] * nv[1] + Rv[0, 2] * nv[2],
                              Rv[1, 0] * nv[0] + Rv[1, 1] * nv[1] + Rv[1, 2] * nv[2],
                              Rv[2, 0] * nv[0] + Rv[2, 1] * nv[1] + Rv[2, 2] * nv[2]], dim=0)
        depth = torch.nan_to_num(allmap[0:1] / alpha, 0.0, 0.0)                                  # depth_ratio = 0
        surf_normal = depth_to_normal(cam, depth).permute(2, 0, 1) * alpha.detach()
        return dict(render=image[:3], specular=image[3:4], roughness=image[4:5], alpha=alpha, normal=normal, depth=depth,
                    surf_normal=surf_normal, dist=allmap[6:7], radii=radii, weight=weight, means2D=means2D, image=image, allmap=allmap)

    # ---- environment pass ------------------------------------------------------------------------------------
    def _prepare_env(self) -> torch.Tensor:
        """Quad corners of the traced set + acceleration structure over them: every training iteration (optix_utils.py:71-85)."""
        with torch.no_grad():
            if self.fused_glue:
                v, f = glue.quad_vertices(self.env_means, self.env_scales, self.env_rotations)
            else:
                v, f = scene.quad_vertices(self.env_means, self.env_scales, self.env_rotations)
            self.tracer.build_acceleration_structure(v, f, rebuild=True)
        return v

    def _prepare_env_async(self, step_start: torch.cuda.Event):
        """The same on a side stream: it depends on the environment parameters only, so it can run underneath the base set's
        rasterizer forward (and the host-side read-backs of a full topology rebuild no longer stall the main stream's queue).
        `step_start` orders it behind everything the main stream did before this step (the previous update of the parameters)."""
        dev = self.env_means.device
        if self._env_stream is None:
            self._env_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self._env_stream):
            self._env_stream.wait_event(step_start)
            v = self._prepare_env()
            done = torch.cuda.Event()
            done.record(self._env_stream)
        v.record_stream(main)
        return v, done

    def trace_env(self, cam: scene.Camera, ref_o: torch.Tensor, ref_d: torch.Tensor, prepared=None) -> dict:
        if prepared is None:
            v = self._prepare_env()
        else:
            v, done = prepared
            torch.cuda.current_stream(v.device).wait_event(done)
        s = SurfelTracingSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, self.env_bg, 1.0, cam.viewmatrix, cam.projmatrix, 3,
                                  cam.campos, False, False, self.max_trace_depth, self.specular_threshold if self.max_trace_depth else 0.0)
        grads3D = torch.zeros_like(self.env_means, requires_grad=True)
        others = torch.cat([self.env_specular, self.env_roughness], dim=-1) if self.max_trace_depth > 0 else None
        rgb, dpt, acc, norm, dist, aux, mid, wet = self.tracer(ref_o, ref_d, v, self.env_means, grads3D, shs=self.env_shs,
                                                               others_precomp=others,
                                                               opacities=self.env_opacities, scales=self.env_scales,
                                                               rotations=self.env_rotations, tracer_settings=s,
                                                               start_from_first=False)
        return dict(rgb=rgb, dpt=dpt, acc=acc, norm=norm, wet=wet, grads3D=grads3D)

    def _ray_mask(self, out: dict) -> torch.Tensor | None:
        if self.ray_filter is None:
            return None
        if self.ray_filter == "specular":
            spec = out["specular"][0]
            return spec > torch.quantile(spec.reshape(-1)[:: max(1, spec.numel() // 4_000_000)], self.specular_filtering_percent)
        if self.ray_filter == "acc":
            return out["alpha"][0] > 0.75
        return out["alpha"][0] > 0.0

    def forward(self, cam: scene.Camera, target: torch.Tensor | None = None) -> dict:
        if self.fused_glue and target is not None and self.ray_filter is None:
            return self._forward_fused(cam, target)
        out = self.rasterize_base(cam)
        rgb = out["render"].permute(1, 2, 0)                     # (H,W,3)
        if self.has_env:
            ray_o, ray_d = scene.camera_rays(cam)
            norm_map = out["normal"].permute(1, 2, 0)
            n = norm_map / (norm_map.norm(dim=-1, keepdim=True) + 1e-8)
            ref_d = ray_d - 2 * (ray_d * n).sum(-1, keepdim=True) * n
            ref_o = ray_o + ray_d * out["depth"].permute(1, 2, 0)
            spec = out["specular"].permute(1, 2, 0)
            msk = self._ray_mask(out)
            if msk is None:
                env = self.trace_env(cam, ref_o.contiguous(), ref_d.contiguous())
                rgb = (1 - spec) * rgb + spec * env["rgb"]      # envgs_sampler.py:474
            else:                                               # compact ray list (1, S, 3), envgs_sampler.py:438-447, :466-470
                env = self.trace_env(cam, ref_o[msk][None].contiguous(), ref_d[msk][None].contiguous())
                blended = rgb.clone()
                blended[msk] = (1 - spec[msk]) * rgb[msk] + spec[msk] * env["rgb"][0]
                rgb = blended
                out["ref_msk"] = msk
            out["env"] = env
        out["rgb"] = rgb
        return out

    def _forward_fused(self, cam: scene.Camera, target: torch.Tensor) -> dict:
        """Same computation with the per-pixel glue in two fused kernels per direction (envgs_b200/glue.py)."""
        overlap = self.has_env and self.async_bvh and self.means.is_cuda
        if overlap:
            step_start = torch.cuda.Event()
            step_start.record()
        image, radii, allmap, weight, means2D = self._raster(cam)
        env = None
        if self.has_env:
            prepared = self._prepare_env_async(step_start) if overlap else None   # queued after the rasterizer forward on the host side
            ref_o, ref_d = glue.reflect_rays(allmap, cam)
            env = self.trace_env(cam, ref_o, ref_d, prepared)
        rgb, loss, parts = glue.blend_loss(image, allmap, None if env is None else env["rgb"], target, cam, self.lambda_normal)
        out = dict(rgb=rgb, loss=loss, loss_parts=parts, radii=radii, weight=weight, means2D=means2D, image=image, allmap=allmap)
        if env is not None:
            out["env"] = env
        return out

    def loss(self, out: dict, target: torch.Tensor) -> torch.Tensor:
        """L1 image loss + 0.04 normal-consistency, so colour, depth, alpha and normal grads are all non-zero."""
        if "loss" in out:
            return out["loss"]
        l1 = (out["rgb"] - target).abs().mean()
        ncons = (1.0 - (out["normal"] * out["surf_normal"]).sum(0)).mean()
        return l1 + self.lambda_normal * ncons

    def flat_grads(self) -> list[torch.Tensor]:
        return [p.grad for p in self.parameters() if p.grad is not None]
