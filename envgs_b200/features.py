"""Fused base-surfel feature op: SH -> view-dependent RGB (+0.5, clamp at 0) with extra columns appended.

Caller-side glue of the rasterizer.  Mirrors what `render()` computes in PyTorch at
easyvolcap/utils/gaussian2d_utils.py:1071-1084: `cat([clamp_min(eval_sh(D, shs, normalize(xyz - campos)) + 0.5, 0), specular, roughness])`.
"""
from __future__ import annotations

import torch

from . import _lib
from . import gradsink


class _SHFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, shs, extra, campos, sh_degree):
        L = _lib.lib()
        dev = means3D.device
        m, s = means3D.detach().float().contiguous(), shs.detach().float().contiguous()
        e = extra.detach().float().contiguous()
        c = campos.detach().float().contiguous()
        P, M, E = m.shape[0], s.shape[1], e.shape[1]
        out = torch.empty((P, 3 + E), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.envgs_sh_features_forward(P, int(sh_degree), M, E, _lib.ptr(m), _lib.ptr(s), _lib.ptr(e), _lib.ptr(c),
                                                   _lib.ptr(out), torch.cuda.current_stream(dev).cuda_stream), "envgs_sh_features_forward")
        ctx.save_for_backward(m, s, c)
        ctx.dims = (P, int(sh_degree), M, E)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _lib.lib()
        m, s, c = ctx.saved_tensors
        P, D, M, E = ctx.dims
        dev = m.device
        g = grad_out.float().contiguous()
        d_means = torch.empty((P, 3), dtype=torch.float32, device=dev)
        d_extra = torch.empty((P, E), dtype=torch.float32, device=dev)
        fac = gradsink.take_factor(s)
        if fac is not None:
            # view-parallel trainer: hand over the rank-one factor (P,3) of this view's SH gradient and the camera centre instead
            # of materialising the (P,M,3) gradient; the trainer all-gathers the factors and envgs_sh_adam_factored rebuilds the sum
            fbuf, cbuf = fac
            cbuf.copy_(c.reshape(-1)[:3])
            with torch.cuda.device(dev):
                _lib.check(L.envgs_sh_features_backward_factored(P, D, M, E, _lib.ptr(m), _lib.ptr(s), _lib.ptr(c), _lib.ptr(g), _lib.ptr(d_means),
                                                                 _lib.ptr(fbuf), _lib.ptr(d_extra), torch.cuda.current_stream(dev).cuda_stream),
                           "envgs_sh_features_backward_factored")
            return d_means, None, d_extra, None, None
        d_shs = gradsink.empty_or_sink(s, (P, M, 3), dtype=torch.float32, device=dev)   # 192 MB at 1 M surfels: written in place
        with torch.cuda.device(dev):
            _lib.check(L.envgs_sh_features_backward(P, D, M, E, _lib.ptr(m), _lib.ptr(s), _lib.ptr(c), _lib.ptr(g), _lib.ptr(d_means),
                                                    _lib.ptr(d_shs), _lib.ptr(d_extra), torch.cuda.current_stream(dev).cuda_stream),
                       "envgs_sh_features_backward")
        return d_means, d_shs, d_extra, None, None


def sh_features(means3D: torch.Tensor, shs: torch.Tensor, extra: torch.Tensor, campos: torch.Tensor, sh_degree: int = 3) -> torch.Tensor:
    """(P,3) means, (P,M,3) SH coefficients, (P,E) extra columns, (3,) camera centre -> (P, 3+E) rasterizer features."""
    return _SHFeatures.apply(means3D, shs, extra, campos, sh_degree)
