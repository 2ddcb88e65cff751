"""Shared test plumbing: run the CUDA operators (through the reference-shaped Python surface, hence through the C ABI)
and the CPU oracle on the same seeded inputs and compare."""
from __future__ import annotations

import math

import torch

from envgs_b200 import scene
from oracle import raster_oracle as RO
from oracle import trace_oracle as TO


def base_features(base: scene.SurfelSet, cam: scene.Camera, C: int) -> torch.Tensor:
    """(P,C) precomputed colours the way EnvGS feeds the C=5 rasterizer (gaussian2d_utils.py:1071-1084)."""
    d = torch.nn.functional.normalize(base.means - cam.campos, dim=-1)
    basis = RO.eval_sh_basis(3, d)
    rgb = torch.clamp_min((basis[:, :, None] * base.shs).sum(1) + 0.5, 0.0)
    if C == 3:
        return rgb.contiguous()
    if C == 5:
        return torch.cat([rgb, base.specular, base.roughness], -1).contiguous()
    extra = torch.cat([base.specular.expand(-1, 3), base.roughness], -1)
    return torch.cat([rgb, extra], -1).contiguous()


def upstream(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def run_cuda_raster(base, cam, C, feats=None, use_sh=False, bg=None, g_color=None, g_others=None, scale_modifier=1.0,
                    device="cuda", transmat=None):
    from envgs_b200.rasterizer import GaussianRasterizationSettings, make_module
    _, _, Rasterizer = make_module(C)
    dev = torch.device(device)
    bg = torch.zeros(3) if bg is None else bg
    s = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg.to(dev), scale_modifier,
                                      cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False, False)
    leaf = lambda t: t.to(dev).clone().requires_grad_(True)
    means, opac = leaf(base.means), leaf(base.opacities)
    scales = rots = cov = None
    if transmat is None:
        scales, rots = leaf(base.scales), leaf(base.rotations)
    else:
        cov = leaf(transmat)
    means2D = torch.zeros_like(means, requires_grad=True)
    shs = colors = None
    if use_sh:
        shs = leaf(base.shs)
    else:
        colors = leaf(feats)
    color, radii, allmap, weight = Rasterizer(s)(means3D=means, means2D=means2D, opacities=opac, shs=shs, colors_precomp=colors,
                                                 scales=scales, rotations=rots, cov3D_precomp=cov)
    out = dict(color=color.detach().cpu(), radii=radii.cpu(), allmap=allmap.detach().cpu(), weight=weight.detach().cpu())
    if g_color is not None:
        loss = (color * g_color.to(dev)).sum() + (allmap * g_others.to(dev)).sum()
        loss.backward()
        out.update(g_means=means.grad.cpu(), g_opac=opac.grad.cpu(), g_means2D=means2D.grad.cpu())
        if transmat is None:
            out.update(g_scales=scales.grad.cpu(), g_rots=rots.grad.cpu())
        else:
            out.update(g_cov=cov.grad.cpu())
        if use_sh:
            out.update(g_shs=shs.grad.cpu())
        else:
            out.update(g_colors=colors.grad.cpu())
    return out


def run_oracle_raster(base, cam, C, feats=None, use_sh=False, bg=None, g_color=None, g_others=None, scale_modifier=1.0,
                      dtype=torch.float64, transmat=None):
    bg = torch.zeros(3) if bg is None else bg
    leaf = lambda t: t.to(dtype).clone().requires_grad_(True)
    means, opac = leaf(base.means), leaf(base.opacities)
    scales = rots = cov = None
    if transmat is None:
        scales, rots = leaf(base.scales), leaf(base.rotations)
    else:
        cov = leaf(transmat)
    shs = colors = None
    if use_sh:
        shs = leaf(base.shs)
    else:
        colors = leaf(feats)
    res = RO.rasterize(means, opac, scales, rots, colors_precomp=colors, shs=shs, sh_degree=3, viewmatrix=cam.viewmatrix,
                       projmatrix=cam.projmatrix, campos=cam.campos, bg=bg, W=cam.W, H=cam.H, scale_modifier=scale_modifier,
                       transMat_precomp=cov, need_T_grad=g_color is not None)
    out = dict(color=res.color.detach().float(), radii=res.radii, allmap=res.allmap.detach().float(),
               weight=res.weight.detach().float(), num_rendered=res.num_rendered)
    if g_color is not None:
        loss = (res.color * g_color.to(dtype)).sum() + (res.allmap * g_others.to(dtype)).sum()
        loss.backward()
        gz = lambda t: torch.zeros_like(t).float() if t.grad is None else t.grad.float()
        out.update(g_means=gz(means), g_opac=gz(opac),
                   g_means2D=RO.densify_signal(res, cam.tanfovx, cam.tanfovy, cam.W, cam.H).float())
        if transmat is None:
            out.update(g_scales=scales.grad.float(), g_rots=rots.grad.float())
        else:
            out.update(g_cov=cov.grad.float())
        if use_sh:
            out.update(g_shs=shs.grad.float())
        else:
            out.update(g_colors=colors.grad.float())
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def trimmed_rel_l2(a: torch.Tensor, b: torch.Tensor, frac: float = 2e-3, min_drop: int = 2) -> float:
    """Relative L2 error over surfels after dropping the `frac` of surfels with the largest error.

    Per-surfel gradients of nearly edge-on / sub-pixel surfels are ill-conditioned in fp32: the fp32 oracle differs from
    the fp64 oracle on exactly those surfels by the same amount the fp32 kernels do ("tolerances"), and a
    single such surfel can dominate the global L2 norm.  The norm in the denominator is the FULL reference norm."""
    a, b = a.detach().double(), b.detach().double()
    d = (a - b).reshape(a.shape[0], -1).norm(dim=1)
    k = max(min_drop, int(frac * d.numel()))
    keep = torch.ones_like(d, dtype=torch.bool)
    if d.numel() > k:
        keep[d.topk(k).indices] = False
    return float((d[keep] ** 2).sum().sqrt() / b.norm().clamp_min(1e-30))


def frac_above(a: torch.Tensor, b: torch.Tensor, tol: float) -> float:
    return float(((a.double() - b.double()).abs() > tol).double().mean())


def image_report(name, a, b, tol=1e-4):
    d = (a.double() - b.double()).abs()
    return f"{name}: max {d.max():.3e} mean {d.mean():.3e} frac>{tol:g} {float((d > tol).double().mean()):.3e}"


def reflect_rays(cam: scene.Camera, dpt: torch.Tensor, norm: torch.Tensor):
    """Reflection rays from a depth map (H,W,1) and a world-space normal map (H,W,3)
    (easyvolcap/models/samplers/envgs_sampler.py:420-431; normalize = x / (|x| + 1e-8), math_utils.py:6-8)."""
    ray_o, ray_d = scene.camera_rays(cam.to(dpt.device))
    n = norm / (norm.norm(dim=-1, keepdim=True) + 1e-8)
    ref_d = ray_d - 2 * (ray_d * n).sum(-1, keepdim=True) * n
    ref_o = ray_o + ray_d * dpt
    return ref_o.contiguous(), ref_d.contiguous()


def synthetic_rays(H, W, seed=5, origin_jitter=0.3):
    """Rays leaving a small region around the origin in all directions (stand-in for reflection rays)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    o = (torch.rand(H, W, 3, generator=g) - 0.5) * 2 * origin_jitter
    d = torch.randn(H, W, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True) * (1.0 + 0.3 * torch.rand(H, W, 1, generator=g))
    return o.float().contiguous(), d.float().contiguous()


def run_cuda_trace(env, cam, ray_o, ray_d, *, use_sh=True, others=False, start_from_first=False, max_trace_depth=0,
                   specular_threshold=0.0, grads=None, device="cuda", training=True, bg=None):
    from envgs_b200.tracer import SurfelTracer, SurfelTracingSettings
    dev = torch.device(device)
    bg = torch.zeros(3) if bg is None else bg
    s = SurfelTracingSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg.to(dev), 1.0, cam.viewmatrix.to(dev),
                              cam.projmatrix.to(dev), 3, cam.campos.to(dev), False, False, max_trace_depth, specular_threshold)
    leaf = lambda t: t.to(dev).clone().requires_grad_(True)
    means, opac, scales, rots = leaf(env.means), leaf(env.opacities), leaf(env.scales), leaf(env.rotations)
    ro, rd = leaf(ray_o), leaf(ray_d)
    shs = colors = oth = None
    if use_sh:
        shs = leaf(env.shs)
    else:
        colors = leaf(torch.clamp_min(env.shs[:, 0, :] * 0.28209479177387814 + 0.5, 0.0))
    if others:
        oth = leaf(torch.cat([env.specular, env.roughness], -1))
    tracer = SurfelTracer().to(dev)
    tracer.train(training)
    v, f = scene.quad_vertices(means.detach(), scales.detach(), rots.detach())
    tracer.build_acceleration_structure(v, f, rebuild=True)
    grads3D = torch.zeros_like(means, requires_grad=True)
    rgb, dpt, acc, norm, dist, aux, mid, wet = tracer(ro, rd, v, means, grads3D, shs=shs, colors_precomp=colors,
                                                      others_precomp=oth, opacities=opac, scales=scales, rotations=rots,
                                                      tracer_settings=s, start_from_first=start_from_first)
    out = dict(rgb=rgb.detach().cpu(), dpt=dpt.detach().cpu(), acc=acc.detach().cpu(), norm=norm.detach().cpu(),
               dist=dist.detach().cpu(), aux=aux.detach().cpu(), mid=mid.detach().cpu(), wet=wet.detach().cpu())
    if grads is not None:
        loss = sum((t * grads[k].to(dev)).sum() for k, t in dict(rgb=rgb, dpt=dpt, acc=acc, norm=norm, aux=aux).items())
        loss.backward()
        out.update(g_means=means.grad.cpu(), g_opac=opac.grad.cpu(), g_scales=scales.grad.cpu(), g_rots=rots.grad.cpu(),
                   g_ray_o=ro.grad.cpu(), g_ray_d=rd.grad.cpu(), g_grads3D=grads3D.grad.cpu())
        if use_sh:
            out.update(g_shs=shs.grad.cpu())
        else:
            out.update(g_colors=colors.grad.cpu())
        if others:
            out.update(g_others=oth.grad.cpu())
    return out


def run_oracle_trace(env, cam, ray_o, ray_d, *, use_sh=True, others=False, start_from_first=False, max_trace_depth=0,
                     specular_threshold=0.0, grads=None, dtype=torch.float64, bg=None, max_hits=96, ray_chunk=2048, device="cpu"):
    """`device="cuda"` only moves the checker's own torch arithmetic (fp64) to the GPU for the large live tests."""
    dev = torch.device(device)
    bg = (torch.zeros(3) if bg is None else bg).to(dev)
    cam = cam.to(dev)
    leaf = lambda t: t.to(device=dev, dtype=dtype).clone().requires_grad_(True)
    means, opac, scales, rots = leaf(env.means), leaf(env.opacities), leaf(env.scales), leaf(env.rotations)
    ro, rd = leaf(ray_o), leaf(ray_d)
    shs = colors = oth = None
    if use_sh:
        shs = leaf(env.shs)
    else:
        colors = leaf(torch.clamp_min(env.shs[:, 0, :] * 0.28209479177387814 + 0.5, 0.0))
    if others:
        oth = leaf(torch.cat([env.specular, env.roughness], -1))
    res = TO.trace(ro, rd, means, opac, scales, rots, shs=shs, colors_precomp=colors, others_precomp=oth, sh_degree=3,
                   campos=cam.campos, bg=bg, start_from_first=start_from_first, max_trace_depth=max_trace_depth,
                   specular_threshold=specular_threshold, projmatrix=cam.projmatrix, max_hits=max_hits, ray_chunk=ray_chunk)
    c = lambda t: t.detach().float().cpu()
    out = dict(rgb=c(res.rgb), dpt=c(res.dpt), acc=c(res.acc), norm=c(res.norm), dist=c(res.dist), aux=c(res.aux),
               mid=c(res.mid), wet=c(res.wet), border=res.border.cpu())
    if grads is not None:
        loss = sum((t * grads[k].to(device=dev, dtype=dtype)).sum()
                   for k, t in dict(rgb=res.rgb, dpt=res.dpt, acc=res.acc, norm=res.norm, aux=res.aux).items())
        loss.backward()
        z = lambda t, ref: torch.zeros_like(ref).float().cpu() if t.grad is None else t.grad.float().cpu()
        out.update(g_means=z(means, means), g_opac=z(opac, opac), g_scales=z(scales, scales), g_rots=z(rots, rots),
                   g_ray_o=z(ro, ro), g_ray_d=z(rd, rd))
        out["g_grads3D"] = out["g_means"] * (env.means - cam.campos.cpu()[None]).norm(dim=-1, keepdim=True) / 2
        if use_sh:
            out.update(g_shs=z(shs, shs))
        else:
            out.update(g_colors=z(colors, colors))
        if others:
            out.update(g_others=z(oth, oth))
    return out


def run_reference_trace(env, cam, ray_o, ray_d, *, use_sh=True, others=False, start_from_first=False, max_trace_depth=0,
                        specular_threshold=0.0, grads=None, device="cuda", training=True, bg=None, backend="emulation"):
    """The UNMODIFIED reference tracer programs on the same inputs as run_cuda_trace, wrapped the way
    diff_surfel_tracing/__init__.py:51-204 wraps them.  backend "emulation": oracle/_ref/ref_trace (OptiX traversal emulated
    by brute force); "emulation_nudge": same with every hit distance moved by one ulp; "optix": the real OptiX runtime
    (oracle/_ref/ref_trace_optix; needs the driver's libnvoptix)."""
    from oracle import ref_trace as RT
    dev = torch.device(device)
    bg = torch.zeros(3) if bg is None else bg
    means, opac, scales, rots = (t.to(dev) for t in (env.means, env.opacities, env.scales, env.rotations))
    shs = colors = oth = None
    if use_sh:
        shs = env.shs.to(dev)
    else:
        colors = torch.clamp_min(env.shs[:, 0, :] * 0.28209479177387814 + 0.5, 0.0).to(dev)
    if others:
        oth = torch.cat([env.specular, env.roughness], -1).to(dev)
    v, f = scene.quad_vertices(means, scales, rots)
    kw = dict(ray_o=ray_o.to(dev), ray_d=ray_d.to(dev), vertices=v, triangles=f, bg=bg.to(dev), means3D=means, shs=shs, degree=3,
              colors_precomp=colors, others_precomp=oth, opacities=opac, scales=scales, scale_modifier=1.0, rotations=rots,
              viewmatrix=cam.viewmatrix.to(dev), projmatrix=cam.projmatrix.to(dev), campos=cam.campos.to(dev),
              start_from_first=start_from_first, max_trace_depth=max_trace_depth, specular_threshold=specular_threshold)
    fwd = RT.trace_forward(training=training, backend=backend, **kw)
    out = {k: fwd[k].cpu() for k in ("rgb", "dpt", "acc", "norm", "aux", "mid", "wet")}
    out["dist"] = fwd["dist"][..., :1].cpu()
    if grads is not None:
        H, W = ray_o.shape[:2]
        up = dict(grads)
        up.setdefault("dist", torch.zeros(H, W, 1))
        bw = RT.trace_backward(fwd, up, backend=backend, **kw)
        out.update(g_means=bw["g_means"].cpu(), g_opac=bw["g_opac"].cpu(), g_scales=bw["g_scales"].cpu(), g_rots=bw["g_rots"].cpu(),
                   g_ray_o=bw["g_ray_o"].cpu(), g_ray_d=bw["g_ray_d"].cpu())
        out["g_grads3D"] = out["g_means"] * (env.means - cam.campos[None]).norm(dim=-1, keepdim=True) / 2
        if use_sh:
            out["g_shs"] = bw["g_shs"].cpu()
        else:
            out["g_colors"] = bw["g_colors"].cpu()
        if others:
            out["g_others"] = bw["g_others"].cpu()
    return out
