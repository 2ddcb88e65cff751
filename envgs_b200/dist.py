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


class ViewParallelTrainer:
    """One optimisation step = every rank renders + differentiates its view, one all-reduce, (optional) fused Adam."""

    def __init__(self, renderer: nn.Module, lr: float = 1e-4, average: bool = True, optimizer: bool = True):
        self.renderer = renderer
        self.flat = FlatParameters(renderer)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.average = average
        self.opt = None
        if optimizer:
            flat_param = nn.Parameter(self.flat.flat, requires_grad=True)
            flat_param.grad = self.flat.flat_grad
            kw = dict(fused=True) if self.flat.flat.is_cuda else {}
            self.opt = torch.optim.Adam([flat_param], lr=lr, eps=1e-15, **kw)
            self._flat_param = flat_param

    def step(self, cam, target) -> torch.Tensor:
        self.flat.zero_grad()
        out = self.renderer(cam, target)
        loss = self.renderer.loss(out, target)
        loss.backward()
        self.flat.collect_stray_grads()
        if self.world > 1:
            dist.all_reduce(self.flat.flat_grad, op=dist.ReduceOp.SUM)   # the single collective of the step
            if self.average:
                self.flat.flat_grad.mul_(1.0 / self.world)
        if self.opt is not None:
            self._flat_param.grad = self.flat.flat_grad
            self.opt.step()
        return loss.detach()
