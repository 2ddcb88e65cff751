"""Fused per-pixel glue between the rasterizer and the tracer: reflected-ray construction and
blend + loss, each one CUDA kernel per direction instead of ~150 eager PyTorch launches per view.  `pipeline.py` keeps the
PyTorch expressions (the reference's way) behind `fused_glue=False`; tests compare the two."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _cam_struct(cam, keep: list) -> "_lib.GlueCamera":
    K, R, c = cam.K.detach().float().contiguous(), cam.R.detach().float().contiguous(), cam.campos.detach().float().contiguous()
    keep += [K, R, c]
    return _lib.GlueCamera(int(cam.H), int(cam.W), float(cam.tanfovx), float(cam.tanfovy), _lib.ptr(K), _lib.ptr(R), _lib.ptr(c))


class _ReflectRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, allmap, cam):
        L = _lib.lib()
        dev = allmap.device
        am = allmap.detach().float().contiguous()
        keep = []
        cs = _cam_struct(cam, keep)
        ref_o = torch.empty((cam.H, cam.W, 3), dtype=torch.float32, device=dev)
        ref_d = torch.empty((cam.H, cam.W, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.envgs_glue_reflect_forward(C.byref(cs), _lib.ptr(am), _lib.ptr(ref_o), _lib.ptr(ref_d),
                                                    torch.cuda.current_stream(dev).cuda_stream), "envgs_glue_reflect_forward")
        ctx.cam = cam
        ctx.save_for_backward(am)
        return ref_o, ref_d

    @staticmethod
    def backward(ctx, g_ro, g_rd):
        L = _lib.lib()
        (am,) = ctx.saved_tensors
        dev = am.device
        cam = ctx.cam
        z = lambda g: torch.zeros((cam.H, cam.W, 3), dtype=torch.float32, device=dev) if g is None else g.float().contiguous()
        g_ro, g_rd = z(g_ro), z(g_rd)
        keep = []
        cs = _cam_struct(cam, keep)
        d_allmap = torch.empty_like(am)
        with torch.cuda.device(dev):
            _lib.check(L.envgs_glue_reflect_backward(C.byref(cs), _lib.ptr(am), _lib.ptr(g_ro), _lib.ptr(g_rd), _lib.ptr(d_allmap),
                                                     torch.cuda.current_stream(dev).cuda_stream), "envgs_glue_reflect_backward")
        return d_allmap, None


class _BlendLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, allmap, env_rgb, target, cam, lambda_n):
        L = _lib.lib()
        dev = image.device
        im, am = image.detach().float().contiguous(), allmap.detach().float().contiguous()
        env = None if env_rgb is None else env_rgb.detach().float().contiguous()
        tg = target.detach().float().contiguous()
        keep = []
        cs = _cam_struct(cam, keep)
        rgb = torch.empty((cam.H, cam.W, 3), dtype=torch.float32, device=dev)
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.envgs_glue_blend_loss_forward(C.byref(cs), _lib.ptr(im), _lib.ptr(am), _lib.ptr(env), _lib.ptr(tg), float(lambda_n),
                                                       _lib.ptr(rgb), _lib.ptr(loss2), torch.cuda.current_stream(dev).cuda_stream),
                       "envgs_glue_blend_loss_forward")
        ctx.cam, ctx.lambda_n, ctx.has_env = cam, float(lambda_n), env is not None
        ctx.save_for_backward(im, am, env if env is not None else torch.empty(0, device=dev), tg, rgb)
        loss = loss2.sum()
        ctx.mark_non_differentiable(loss2)
        return rgb, loss, loss2

    @staticmethod
    def backward(ctx, g_rgb, g_loss, _g_loss2):
        L = _lib.lib()
        im, am, env, tg, rgb = ctx.saved_tensors
        dev = im.device
        cam = ctx.cam
        env_p = env if ctx.has_env else None
        gl = (torch.zeros(1, dtype=torch.float32, device=dev) if g_loss is None else g_loss.float().reshape(1).contiguous())
        grgb = None if g_rgb is None else g_rgb.float().contiguous()
        keep = []
        cs = _cam_struct(cam, keep)
        d_image, d_allmap = torch.empty_like(im), torch.empty_like(am)
        d_env = torch.empty_like(rgb) if ctx.has_env else None
        with torch.cuda.device(dev):
            _lib.check(L.envgs_glue_blend_loss_backward(C.byref(cs), _lib.ptr(im), _lib.ptr(am), _lib.ptr(env_p), _lib.ptr(tg), _lib.ptr(rgb),
                                                        ctx.lambda_n, _lib.ptr(gl), _lib.ptr(grgb), _lib.ptr(d_image), _lib.ptr(d_allmap),
                                                        _lib.ptr(d_env), torch.cuda.current_stream(dev).cuda_stream),
                       "envgs_glue_blend_loss_backward")
        return d_image, d_allmap, d_env, None, None, None


def reflect_rays(allmap: torch.Tensor, cam):
    """allmap (7,H,W) -> (ref_o, ref_d) each (H,W,3)."""
    return _ReflectRays.apply(allmap, cam)


def blend_loss(image: torch.Tensor, allmap: torch.Tensor, env_rgb, target: torch.Tensor, cam, lambda_n: float = 0.04):
    """image (5,H,W), allmap (7,H,W), env rgb (H,W,3) or None, target (H,W,3) -> (rgb (H,W,3), loss scalar, loss parts (2,))."""
    return _BlendLoss.apply(image, allmap, env_rgb, target, cam, lambda_n)


_FACES: dict = {}


def quad_vertices(means: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor):
    """Fused scene.quad_vertices (no gradient, like the reference's GAS build): (4P,3) corners + the cached (2P,3) faces."""
    P = means.shape[0]
    m, s, r = (t.detach().to(torch.float32).contiguous() for t in (means, scales, rotations))
    v = torch.empty((4 * P, 3), dtype=torch.float32, device=means.device)
    with torch.cuda.device(means.device):
        _lib.check(_lib.lib().envgs_glue_quad_vertices(P, _lib.ptr(m), _lib.ptr(s), _lib.ptr(r), _lib.ptr(v),
                                                       torch.cuda.current_stream(means.device).cuda_stream), "envgs_glue_quad_vertices")
    key = (P, str(means.device))
    f = _FACES.get(key)
    if f is None:
        idx = torch.arange(4 * P, device=means.device).reshape(-1, 4)
        f = torch.stack([idx[:, :3], idx[:, 1:]], dim=1).reshape(-1, 3).int().contiguous()
        _FACES.clear()                                   # one face table at a time (P changes only at densification)
        _FACES[key] = f
    return v, f

