"""View-parallel training step: one process per GPU, every rank holds a replica of both surfel sets and renders its own
view(s); gradients are summed with ONE NCCL all-reduce over a single flat fp32 buffer.

The reference has no usable multi-GPU path for EnvGS (its DDP wrapper breaks when densification re-creates parameters,
SURVEY.md §2.4), so this is new host logic: parameters are views into one flat tensor and their `.grad`s are views into one
flat gradient tensor, which the backward kernels' outputs accumulate into — no gather copy before the collective.
Works with backend `nccl` (GPU) and `gloo` (CPU tests, world_size 2).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn


class FlatParameters:
    """Re-homes every parameter of `module` (and its gradient) into two contiguous fp32 buffers."""

    def __init__(self, module: nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            self.offsets.append((off, n))
            off += n
        self.total = total

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, (off, n) in zip(self.params, self.offsets):   # autograd may have replaced the view; re-attach
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + n].view_as(p)

    def collect_stray_grads(self):
        """If autograd swapped a .grad tensor instead of accumulating in place, fold it back into the flat buffer."""
        for p, (off, n) in zip(self.params, self.offsets):
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                self.flat_grad[off:off + n].add_(p.grad.reshape(-1))
                p.grad = self.flat_grad[off:off + n].view_as(p)


def init_process_group_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; single process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_views(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin view assignment: rank r renders views r, r+world, ...."""
    return list(range(rank, n_views, world))


class FlatAdam:
    """Adam over the flat buffers in ONE kernel (`envgs_adam_step`): gradient scaling (the 1/world average), moment and
    parameter update and the zero-fill of the gradient buffer for the next step.  CUDA only; arithmetic of
    torch.optim.Adam(fused=True) without weight decay (the reference: EVC/runners/optimizers.py:163-201)."""

    def __init__(self, flat: torch.Tensor, flat_grad: torch.Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-15):
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat_grad.is_contiguous()
        self.flat, self.flat_grad, self.lr, self.betas, self.eps = flat, flat_grad, lr, betas, eps
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        self.t = 0

    def step(self, grad_scale: float = 1.0, zero_grad: bool = True) -> None:
        from . import _lib
        self.t += 1
        with torch.cuda.device(self.flat.device):
            _lib.check(_lib.lib().envgs_adam_step(self.flat.numel(), _lib.ptr(self.flat), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                                  _lib.ptr(self.exp_avg_sq), self.t, self.lr, self.betas[0], self.betas[1], self.eps,
                                                  grad_scale, int(zero_grad), torch.cuda.current_stream(self.flat.device).cuda_stream),
                       "envgs_adam_step")


class ViewParallelTrainer:
    """One optimisation step = every rank renders + differentiates its view, one all-reduce, one fused Adam/scale/zero pass."""

    def __init__(self, renderer: nn.Module, lr: float = 1e-4, average: bool = True, optimizer: bool = True):
        self.renderer = renderer
        self.flat = FlatParameters(renderer)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.average = average
        self.opt = None
        self.fused = None
        self._grad_is_zero = False
        if optimizer and self.flat.flat.is_cuda:
            self.fused = FlatAdam(self.flat.flat, self.flat.flat_grad, lr=lr)
        elif optimizer:                                  # CPU (gloo tests): torch's Adam over the same flat buffers
            flat_param = nn.Parameter(self.flat.flat, requires_grad=True)
            flat_param.grad = self.flat.flat_grad
            self.opt = torch.optim.Adam([flat_param], lr=lr, eps=1e-15)
            self._flat_param = flat_param

    def step(self, cam, target) -> torch.Tensor:
        if not self._grad_is_zero:                       # the fused update leaves the gradient buffer zeroed for the next step
            self.flat.zero_grad()
        out = self.renderer(cam, target)
        loss = self.renderer.loss(out, target)
        loss.backward()
        self.flat.collect_stray_grads()
        if self.world > 1:
            dist.all_reduce(self.flat.flat_grad, op=dist.ReduceOp.SUM)   # the single collective of the step
        scale = 1.0 / self.world if (self.world > 1 and self.average) else 1.0
        if self.fused is not None:
            self.fused.step(grad_scale=scale, zero_grad=True)
            self._grad_is_zero = True
        else:
            if scale != 1.0:
                self.flat.flat_grad.mul_(scale)
            if self.opt is not None:
                self._flat_param.grad = self.flat.flat_grad
                self.opt.step()
        return loss.detach()
