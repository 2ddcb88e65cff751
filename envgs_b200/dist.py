"""View-parallel training step: one process per GPU, every rank holds a replica of both surfel sets and renders its own
view(s); gradients are summed over NCCL in three groups of ONE flat fp32 buffer: the
environment set from a gradient hook underneath the rasterizer's backward, the base SH gradient as rank-one factors, the rest
of the base set as reduce-scatter -> 1/N-sharded fused Adam -> in-place all-gather.

The reference has no usable multi-GPU path for EnvGS (its DDP wrapper breaks when densification re-creates parameters,
SURVEY.md §2.4), so this is new host logic: parameters are views into one flat tensor and their `.grad`s are views into one
flat gradient tensor, which the backward kernels write into directly (gradsink.py) — no gather copy before a collective.
Works with backend `nccl` (GPU) and `gloo` (CPU tests, world_size 2: one all-reduce per group, torch Adam).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import gradsink


class FlatParameters:
    """Re-homes every parameter of `module` (and its gradient) into two contiguous fp32 buffers.

    Parameters whose name starts with one of `early_prefixes` form the EARLY group at the tail of the buffers: their gradients
    are complete long before the end of the backward pass (the environment set: the tracer's backward runs before the
    rasterizer's), so their exchange can start while the rest is still being computed.  Both groups are padded to a multiple
    of `pad` floats so that a group can be reduce-scattered across any world size that divides `pad`."""

    def __init__(self, module: nn.Module, early_prefixes: tuple = ("env_",), pad: int = 1024, factored_names: tuple = ()):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        fact = [(n, p) for n, p in named if n in factored_names]
        late = [(n, p) for n, p in named if not n.startswith(tuple(early_prefixes)) and n not in factored_names]
        early = [(n, p) for n, p in named if n.startswith(tuple(early_prefixes)) and n not in factored_names]
        self.names = [n for n, _ in late + fact + early]
        self.params = [p for _, p in late + fact + early]
        rnd = lambda x: (x + pad - 1) // pad * pad
        n_late, n_fact, n_early = (sum(p.numel() for _, p in grp) for grp in (late, fact, early))
        self.late_range = (0, rnd(n_late))
        self.factored_range = (rnd(n_late), rnd(n_late) + rnd(n_fact))      # the base SH coefficients when their exchange is factored
        self.early_range = (self.factored_range[1], self.factored_range[1] + rnd(n_early))
        self.n_early_params = len(early)
        self._group_starts = {len(late): self.factored_range[0], len(late) + len(fact): self.early_range[0]}
        total = self.early_range[1]
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for k, p in enumerate(self.params):
            off = self._group_starts.get(k, off)
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            self.offsets.append((off, n))
            off += n
        self.total = total

    def view_of(self, k: int, buf: torch.Tensor) -> torch.Tensor:
        off, n = self.offsets[k]
        return buf[off:off + n].view_as(self.params[k])

    def grad_is_view(self, k: int) -> bool:
        p, (off, _) = self.params[k], self.offsets[k]
        return p.grad is not None and p.grad.data_ptr() == self.flat_grad.data_ptr() + 4 * off

    def register_sinks(self) -> None:
        """Let the operators' backward kernels write these parameters' gradients straight into `flat_grad` (gradsink.py)."""
        for k, p in enumerate(self.params):
            gradsink.register(p, self.view_of(k, self.flat_grad))

    def begin_backward(self) -> None:
        """`.grad = None` so that autograd ADOPTS the first gradient tensor of each parameter instead of adding it to an
        existing one — for a parameter with a sink that tensor is the flat-buffer slice the kernel already wrote."""
        gradsink.begin_step()
        for p in self.params:
            p.grad = None

    def home_grad(self, k: int) -> None:
        """After backward: make sure parameter k's complete gradient sits in its flat-buffer slice and `.grad` views it."""
        p = self.params[k]
        if p.grad is not None and not self.grad_is_view(k):     # autograd built a separate tensor (no sink, or several contributions)
            self.view_of(k, self.flat_grad).copy_(p.grad)       # ... which holds the COMPLETE gradient: overwrite, do not add
        p.grad = self.view_of(k, self.flat_grad)

    def zero_grad(self):
        self.flat_grad.zero_()
        for k in range(len(self.params)):
            self.params[k].grad = self.view_of(k, self.flat_grad)

    def collect_stray_grads(self, start: int = 0):
        for k in range(start, len(self.params)):
            self.home_grad(k)


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
    """One optimisation step = every rank renders + differentiates its view, the gradients are summed across ranks, Adam runs.

    Exchange schedule (N > 1):
      * EARLY group (environment set, ~20 % of the bytes): one asynchronous all-reduce launched from a gradient hook the
        moment the tracer's backward has produced the last of its gradients — it runs on NCCL's stream underneath the
        rasterizer's backward kernels (~3 ms of cover for ~0.2 ms of traffic);
      * LATE group (base set): complete only when the backward pass ends.  With NCCL: reduce-scatter -> Adam on this rank's
        1/N shard (moments are stored for the shard only) -> all-gather of the updated parameters, so the optimizer pass and its
        state shrink by N; the collective bytes equal one all-reduce.  Other backends (gloo CPU tests): one all-reduce.
    The gradient average (1/N) is folded into the fused Adam kernel."""

    def __init__(self, renderer: nn.Module, lr: float = 1e-4, average: bool = True, optimizer: bool = True, overlap: bool = True,
                 shard_optimizer: bool = True, factored_sh: bool = True):
        self.renderer = renderer
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        cuda0 = next(renderer.parameters()).is_cuda
        # Factored exchange of the base SH gradient (features.cu, sh_adam_factored_kernel): needs the fused SH feature op
        # (it produces the rank-one factor), NCCL and an optimizer to fuse the rebuild into.
        self.factored = bool(factored_sh and optimizer and cuda0 and self.world > 1 and dist.get_backend() == "nccl"
                             and getattr(renderer, "fused_features", False) and hasattr(renderer, "shs"))
        self.flat = FlatParameters(renderer, pad=1024 * 8, factored_names=("shs",) if self.factored else ())
        self.flat.register_sinks()
        self.average = average
        self.opt = None
        self.fused = self.fused_early = None
        self._grad_is_zero = False
        self.loss_ready = None
        self._early_work = None
        self._early_seen = 0
        cuda = self.flat.flat.is_cuda
        self.overlap = overlap and self.world > 1 and self.flat.n_early_params > 0
        self.shard = (shard_optimizer and optimizer and cuda and self.world > 1 and dist.get_backend() == "nccl"
                      and (self.flat.late_range[1] - self.flat.late_range[0]) % self.world == 0)
        l0, l1 = self.flat.late_range
        e0, e1 = self.flat.early_range
        self.fused_sh = None
        if self.factored:
            f0, f1 = self.flat.factored_range
            dev = self.flat.flat.device
            shs = renderer.shs
            P, M = shs.shape[0], shs.shape[1]
            self._factor_local = torch.zeros(P, 3, dtype=torch.float32, device=dev)
            # The factor all-gather is the largest exposed collective (12 B x P x N received per rank) and the SH update that consumes
            # it is the largest optimizer kernel: split both into surfel chunks, so chunk k's update runs under chunk k+1's gather.
            n_chunks = max(1, int(os.environ.get("ENVGS_SH_EXCHANGE_CHUNKS", 4 if P >= (1 << 18) else 1)))
            per = -(-P // n_chunks)
            self._sh_chunks = [(a, min(a + per, P)) for a in range(0, P, per)]
            self._factors_all = [torch.zeros(self.world, b - a, 3, dtype=torch.float32, device=dev) for a, b in self._sh_chunks]
            self._campos_local = torch.zeros(3, dtype=torch.float32, device=dev)
            self._campos_all = torch.zeros(self.world, 3, dtype=torch.float32, device=dev)
            self.fused_sh = FlatAdam(self.flat.flat[f0:f1], self.flat.flat_grad[f0:f1], lr=lr)   # owns the SH moments and step count
            gradsink.register_factor(shs, self._factor_local, self._campos_local)
        if optimizer and cuda:
            if self.shard:
                n = (l1 - l0) // self.world
                self._shard_slice = slice(l0 + self.rank * n, l0 + (self.rank + 1) * n)
                self._grad_shard = torch.zeros(n, dtype=torch.float32, device=self.flat.flat.device)
                self.fused = FlatAdam(self.flat.flat[self._shard_slice], self._grad_shard, lr=lr)
                self.fused_early = FlatAdam(self.flat.flat[e0:e1], self.flat.flat_grad[e0:e1], lr=lr) if e1 > e0 else None
            elif self.factored:
                self.fused = FlatAdam(self.flat.flat[l0:l1], self.flat.flat_grad[l0:l1], lr=lr)
                self.fused_early = FlatAdam(self.flat.flat[e0:e1], self.flat.flat_grad[e0:e1], lr=lr) if e1 > e0 else None
            else:
                self.fused = FlatAdam(self.flat.flat, self.flat.flat_grad, lr=lr)
        elif optimizer:                                  # CPU (gloo tests): torch's Adam over the same flat buffers
            flat_param = nn.Parameter(self.flat.flat, requires_grad=True)
            flat_param.grad = self.flat.flat_grad
            self.opt = torch.optim.Adam([flat_param], lr=lr, eps=1e-15)
            self._flat_param = flat_param
        if self.overlap:
            first_early = len(self.flat.params) - self.flat.n_early_params
            for k in range(first_early, len(self.flat.params)):
                self.flat.params[k].register_post_accumulate_grad_hook(self._make_hook(k))

    def _make_hook(self, k):
        def hook(_param):
            self._early_seen += 1
            if self._early_seen == self.flat.n_early_params and self._early_work is None:
                first_early = len(self.flat.params) - self.flat.n_early_params
                for j in range(first_early, len(self.flat.params)):      # every early gradient is final here (post-accumulate)
                    self.flat.home_grad(j)
                if True:
                    e0, e1 = self.flat.early_range
                    # NCCL's stream picks up after everything queued on the compute stream so far (the tracer's backward) and
                    # runs concurrently with what is queued next (the rasterizer's backward)
                    self._early_work = dist.all_reduce(self.flat.flat_grad[e0:e1], op=dist.ReduceOp.SUM, async_op=True)
        return hook

    def step(self, cam, target, loss_out: torch.Tensor | None = None) -> torch.Tensor:
        """`loss_out` (pinned host tensor, optional): the loss is copied into it as soon as the forward pass has produced it and
        `self.loss_ready` (a CUDA event) marks the copy — a caller that logs the loss every step waits on that event instead of
        on the whole step, so the host keeps queueing the next step while the backward pass and the update still run."""
        if not self._grad_is_zero:                       # the fused update leaves the gradient buffer zeroed for the next step
            self.flat.zero_grad()
        self._early_seen, self._early_work = 0, None
        out = self.renderer(cam, target)
        loss = self.renderer.loss(out, target)
        if loss_out is not None:
            loss_out.copy_(loss.detach().reshape(loss_out.shape), non_blocking=True)
            if loss.is_cuda:
                if self.loss_ready is None:
                    self.loss_ready = torch.cuda.Event()
                self.loss_ready.record()
        self.flat.begin_backward()
        loss.backward()
        scale = 1.0 / self.world if (self.world > 1 and self.average) else 1.0
        l0, l1 = self.flat.late_range
        e0, e1 = self.flat.early_range
        f0, f1 = self.flat.factored_range
        sh_factored = False
        sh_works, works = [], []
        if self.world > 1:
            self.flat.collect_stray_grads(0)             # every gradient home in the flat buffer (early group: no-op after the hook)
            if self.factored:                            # issued first: the SH update below only waits for these
                sh_factored = gradsink.factor_written(self.renderer.shs)
                if sh_factored:                          # 12 B per surfel and rank instead of 192 B both ways
                    sh_works.append(dist.all_gather_into_tensor(self._campos_all, self._campos_local, async_op=True))
                    for (a, b), buf in zip(self._sh_chunks, self._factors_all):
                        sh_works.append(dist.all_gather_into_tensor(buf, self._factor_local[a:b], async_op=True))
                else:                                    # the SH gradient arrived the ordinary way (no fused feature op in this graph)
                    sh_works.append(dist.all_reduce(self.flat.flat_grad[f0:f1], op=dist.ReduceOp.SUM, async_op=True))
            if self._early_work is None and e1 > e0:     # no hook launch (overlap off): exchange the early group now
                works.append(dist.all_reduce(self.flat.flat_grad[e0:e1], op=dist.ReduceOp.SUM, async_op=True))
            if self.shard:
                works.append(dist.reduce_scatter_tensor(self._grad_shard, self.flat.flat_grad[l0:l1], op=dist.ReduceOp.SUM, async_op=True))
            elif l1 > l0:
                works.append(dist.all_reduce(self.flat.flat_grad[l0:l1], op=dist.ReduceOp.SUM, async_op=True))
            if self._early_work is not None:
                works.append(self._early_work)
            if self.fused is None:                       # no fused optimizer: everything has to be home before the generic path
                sh_works, works = [], sh_works + works
                for w in works:
                    w.wait()
                works = []
        else:
            self.flat.collect_stray_grads()
        if self.fused is not None:
            if self.fused_sh is not None:                # before the positions move: the basis is a function of (mean - campos)
                if sh_factored:                          # runs while the later chunks and the base set's reduce-scatter are still on NCCL's stream
                    from . import _lib
                    shs, a = self.renderer.shs, self.fused_sh
                    a.t += 1
                    row = shs.shape[1] * 3
                    flat_sh = self.flat.flat[f0:f1]
                    sh_works[0].wait()
                    with torch.cuda.device(shs.device):
                        for (c0, c1), buf, w in zip(self._sh_chunks, self._factors_all, sh_works[1:]):
                            w.wait()
                            _lib.check(_lib.lib().envgs_sh_adam_factored(
                                c1 - c0, 3, shs.shape[1], self.world, _lib.ptr(self.renderer.means[c0:c1]), _lib.ptr(self._campos_all),
                                _lib.ptr(buf), _lib.ptr(flat_sh[c0 * row:c1 * row]), _lib.ptr(a.exp_avg[c0 * row:c1 * row]),
                                _lib.ptr(a.exp_avg_sq[c0 * row:c1 * row]), a.t, a.lr, a.betas[0], a.betas[1], a.eps, scale,
                                torch.cuda.current_stream(shs.device).cuda_stream), "envgs_sh_adam_factored")
                else:
                    for w in sh_works:
                        w.wait()
                    self.fused_sh.step(grad_scale=scale, zero_grad=True)
            else:
                for w in sh_works:
                    w.wait()
            for w in works:
                w.wait()
            if self.shard:
                self.fused.step(grad_scale=scale, zero_grad=False)
                # in place: rank r's slice is already home.  The early group's update and the zero-fill run underneath it.
                gather = dist.all_gather_into_tensor(self.flat.flat[l0:l1], self.flat.flat[self._shard_slice], async_op=True)
                if self.fused_early is not None:
                    self.fused_early.step(grad_scale=scale, zero_grad=True)
                self.flat.flat_grad[l0:l1].zero_()
                gather.wait()
            else:
                self.fused.step(grad_scale=scale, zero_grad=True)
                if self.fused_early is not None and self.factored:
                    self.fused_early.step(grad_scale=scale, zero_grad=True)
            self._grad_is_zero = True
        else:
            if scale != 1.0:
                self.flat.flat_grad.mul_(scale)
            if self.opt is not None:
                self._flat_param.grad = self.flat.flat_grad
                self.opt.step()
        return loss.detach()
