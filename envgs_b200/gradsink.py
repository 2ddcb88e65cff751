"""Gradient sinks: let the operators' backward kernels write parameter gradients straight into the trainer's flat gradient
buffer ("backward kernels write straight into this buffer").

Without a sink, every backward allocates a zero-filled tensor per gradient, the kernel fills it, and autograd's AccumulateGrad
then adds it into the parameter's `.grad` (a view of the flat buffer): one extra read + read + write pass over ~300 MB per step
plus the zero fills.  With a sink registered for a parameter, the backward wrapper asks `take(saved_tensor)` for its output
buffer and gets the parameter's slice of the flat gradient buffer itself (the trainer keeps it zeroed between steps: the
fused Adam pass clears what it consumed); the wrapper returns that very tensor as the gradient and, because the trainer leaves
`param.grad = None` before `backward()`, AccumulateGrad adopts it without a copy.  A second gradient for the same parameter in
the same step (e.g. the base means receive one from the rasterizer and one from the SH feature kernel) falls back to a fresh
tensor and is accumulated by autograd as usual.
"""
from __future__ import annotations

import torch

_SINKS: dict = {}     # (device index, data_ptr of the parameter) -> [flat view, taken-this-step flag]


def register(param: torch.Tensor, grad_view: torch.Tensor) -> None:
    assert grad_view.numel() == param.numel() and grad_view.dtype == torch.float32
    _SINKS[(param.device.index, param.data_ptr())] = [grad_view, False]


def clear() -> None:
    _SINKS.clear()
    _FACTORS.clear()


def begin_step() -> None:
    for v in _SINKS.values():
        v[1] = False
    begin_step_factors()


def take(saved: torch.Tensor, shape) -> torch.Tensor | None:
    """The sink of the parameter `saved` aliases (autograd hands the backward a detached alias: same storage, same pointer), viewed
    as `shape`; None if there is none or it was already handed out in this step."""
    if not _SINKS:
        return None
    ent = _SINKS.get((saved.device.index, saved.data_ptr()))
    if ent is None or ent[1] or ent[0].numel() != saved.numel():
        return None
    ent[1] = True
    return ent[0].view(shape)


def zeros_or_sink(saved: torch.Tensor, shape, **kw) -> torch.Tensor:
    """Buffer for a gradient the kernel ACCUMULATES into (atomics / += on a zeroed buffer)."""
    t = take(saved, shape)
    return t if t is not None else torch.zeros(shape, **kw)


def empty_or_sink(saved: torch.Tensor, shape, **kw) -> torch.Tensor:
    """Buffer for a gradient the kernel overwrites completely."""
    t = take(saved, shape)
    return t if t is not None else torch.empty(shape, **kw)


# ---- factored SH gradient (view-parallel training, dist.py) -----------------------------------------------------------------
_FACTORS: dict = {}   # (device index, data_ptr of the SH parameter) -> [(P,3) buffer, camera-centre slot (3,), written-this-step flag]


def register_factor(sh_param: torch.Tensor, factor_buf: torch.Tensor, campos_buf: torch.Tensor) -> None:
    _FACTORS[(sh_param.device.index, sh_param.data_ptr())] = [factor_buf, campos_buf, False]


def take_factor(saved_shs: torch.Tensor):
    """If the trainer asked for the rank-one factor of this SH parameter's gradient instead of the gradient itself: the (P,3)
    buffer to write it into and the slot that receives this view's camera centre; else None."""
    ent = _FACTORS.get((saved_shs.device.index, saved_shs.data_ptr())) if _FACTORS else None
    if ent is None or ent[2]:
        return None
    ent[2] = True
    return ent[0], ent[1]


def begin_step_factors() -> None:
    for v in _FACTORS.values():
        v[2] = False


def factor_written(sh_param: torch.Tensor) -> bool:
    ent = _FACTORS.get((sh_param.device.index, sh_param.data_ptr()))
    return bool(ent and ent[2])
