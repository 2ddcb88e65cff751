"""Python operator surface of the surfel rasterizer, mirroring the reference packages
`diff_surfel_rasterization_wet[_ch05|_ch07]` (RAST/diff_surfel_rasterization_wet/__init__.py:21-222): same class
names, argument order, defaults, return arity, exceptions and autograd contract.  The arithmetic lives in
libenvgs_b200.so and is reached through the C ABI (`envgs_raster_forward/backward`, include/envgs_b200.h).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from . import gradsink


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def _f32c(t: torch.Tensor | None, device) -> torch.Tensor | None:
    if t is None:
        return None
    if t.device != device:
        raise RuntimeError("all rasterizer inputs must be CUDA tensors on the same device")
    return t.detach().to(torch.float32).contiguous()


_ALLOC_QUANTUM = 32 << 20
_ALLOC_HWM: dict = {}   # (device, H, W, C) -> largest rounded `binning` request so far


def reset_scratch_high_water_marks() -> None:
    """Forget the binning-buffer high-water marks (call after densify / prune shrank the set, or when switching to a much
    smaller resolution); the next frames re-learn them."""
    _ALLOC_HWM.clear()
    _BUDGETS.clear()


# ---------------------------------------------------------------------------------------------------------------------------------
# Host-sync-free binning.  The reference blocks the host once per rasterizer call to read the (tile, surfel) instance count that
# sizes the sort (rasterizer_impl.cu:283).  With the mode on, the first call for a given (device, H, W, C, P) still does that, and
# every later one hands the library a capacity (`margin` x the largest count seen so far) instead: the count stays on the device,
# is copied to pinned memory asynchronously and only looked at — without waiting — before a later call.  A frame whose count
# exceeded the capacity lost its excess instances; that is reported with a warning and the budget grows.
# ---------------------------------------------------------------------------------------------------------------------------------
_SYNC_FREE = {"enabled": os.environ.get("ENVGS_RASTER_SYNC_FREE", "0") == "1", "margin": 1.3}
_CAP_QUANTUM = 1 << 16
_BUDGETS: dict = {}


def set_sync_free(enabled: bool, margin: float = 1.3) -> None:
    _SYNC_FREE["enabled"] = bool(enabled)
    _SYNC_FREE["margin"] = float(margin)
    _BUDGETS.clear()


class _InstanceBudget:
    SLOTS = 32   # pinned read-back slots, reused round-robin: no host allocation (cudaHostAlloc synchronises) on the step path

    def __init__(self):
        self.max_seen = -1
        self.host = None
        self.events = [None] * self.SLOTS
        self.pending = []     # slot indices in issue order
        self.head = 0
        self.overflows = 0

    def poll(self):
        while self.pending and self.events[self.pending[0]].query():
            slot = self.pending.pop(0)
            count, over = int(self.host[slot, 0]), int(self.host[slot, 1])
            if over:
                self.overflows += 1
                warnings.warn(f"rasterizer: a frame produced {count} (tile, surfel) instances, more than the sync-free capacity; "
                              "the excess was dropped from that frame and the capacity grows", RuntimeWarning)
            self.max_seen = max(self.max_seen, count)

    def capacity(self):
        if self.max_seen < 0:
            return 0
        want = max(int(math.ceil(self.max_seen * _SYNC_FREE["margin"])), 1)
        return (want + _CAP_QUANTUM - 1) // _CAP_QUANTUM * _CAP_QUANTUM

    def watch(self, count_dev):
        if self.host is None:
            self.host = torch.zeros(self.SLOTS, 2, dtype=torch.int32).pin_memory()
        slot = self.head
        if slot in self.pending:          # the host is a full ring ahead of the device: this frame's count goes unobserved
            return
        self.head = (slot + 1) % self.SLOTS
        self.host[slot].copy_(count_dev, non_blocking=True)
        if self.events[slot] is None:
            self.events[slot] = torch.cuda.Event()
        self.events[slot].record(torch.cuda.current_stream(count_dev.device))
        self.pending.append(slot)


class _Scratch:
    """Allocator callbacks handed to the C ABI: three torch byte tensors owned by the autograd context."""

    def __init__(self, device, shape_key=None):
        self.device = device
        self.shape_key = shape_key
        self.bufs = {}
        self.cbs = {}
        for name in ("geom", "binning", "img"):
            self.cbs[name] = _lib.ALLOC_FN(self._make(name))

    def _make(self, name):
        def alloc(_user, nbytes):
            # The binning buffer scales with the number of (tile, surfel) instances, which changes from view to view.  A
            # varying request size makes torch's caching allocator split and re-carve its big blocks step after step
            # (measured: reserved memory creeping from 3 to 10 GB and a 30-110 ms cudaMalloc stall every few views).  Sizes
            # are therefore rounded up to a coarse quantum and never shrink below the largest request seen on this device:
            # after the first few views every step issues the identical allocation sequence and the allocator is quiet.
            # Only `binning` varies; `geom` and `img` are fixed for a given (P, H, W) and are allocated at their exact size, so a
            # large frame does not inflate the scratch of later low-resolution / other-channel calls.
            n = max(int(nbytes), 1)
            if name == "binning" and n > (4 << 20):
                n = (n + _ALLOC_QUANTUM - 1) // _ALLOC_QUANTUM * _ALLOC_QUANTUM
                key = (str(self.device), self.shape_key)
                n = max(n, _ALLOC_HWM.get(key, 0))
                _ALLOC_HWM[key] = n
            try:
                t = torch.empty(n, dtype=torch.uint8, device=self.device)
            except Exception:  # surfaced by the library as ENVGS_ERR_ALLOC
                return None
            self.bufs[name] = t
            return t.data_ptr()
        return alloc

    def get(self, name):
        t = self.bufs.get(name)
        return t if t is not None else torch.empty(0, dtype=torch.uint8, device=self.device)

    def close(self):
        """Break the self -> callback thunk -> closure -> self cycle: otherwise the ~300 MB of scratch this object holds
        is only released when Python's cyclic GC happens to run, and the caching allocator keeps growing (3 -> 10 GB)."""
        self.bufs = {}
        self.cbs = {}


def _make_args(C_channels, settings, bg, means3D, sh, colors, opacities, scales, rotations, cov, viewmatrix,
               projmatrix, campos):
    a = _lib.RasterArgs()
    a.P = means3D.shape[0]
    a.C = C_channels
    a.sh_degree = int(settings.sh_degree)
    a.sh_coeffs = sh.shape[1] if sh.numel() else 0
    a.width = int(settings.image_width)
    a.height = int(settings.image_height)
    a.tan_fovx = float(settings.tanfovx)
    a.tan_fovy = float(settings.tanfovy)
    a.scale_modifier = float(settings.scale_modifier)
    a.prefiltered = int(bool(settings.prefiltered))
    a.debug = int(bool(settings.debug))
    a.background = _lib.ptr(bg)
    a.means3D = _lib.ptr(means3D)
    a.shs = _lib.ptr(sh)
    a.colors_precomp = _lib.ptr(colors)
    a.opacities = _lib.ptr(opacities)
    a.scales = _lib.ptr(scales)
    a.rotations = _lib.ptr(rotations)
    a.transMat_precomp = _lib.ptr(cov)
    a.viewmatrix = _lib.ptr(viewmatrix)
    a.projmatrix = _lib.ptr(projmatrix)
    a.campos = _lib.ptr(campos)
    return a


def _pad_bg(bg: torch.Tensor, C_channels: int, device) -> torch.Tensor:
    # EnvGS hands a 3-float background to the 5-channel kernels (gaussian2d_sampler.py:173-174), which the
    # reference then reads out of bounds (forward.cu:431).  Extra channels are zero here.
    bg = bg.detach().to(device=device, dtype=torch.float32).reshape(-1)
    if bg.numel() < C_channels:
        bg = torch.cat([bg, bg.new_zeros(C_channels - bg.numel())])
    return bg[:C_channels].contiguous()


def make_module(C_channels: int):
    """Build the (function, Function, Module) triple for one NUM_CHANNELS variant."""

    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings):
            L = _lib.lib()
            if means3D.dim() != 2 or means3D.shape[1] != 3:
                raise RuntimeError("means3D must have dimensions (num_points, 3)")
            if not means3D.is_cuda:
                raise RuntimeError("means3D must be a CUDA tensor")
            device = means3D.device
            s = raster_settings
            P = means3D.shape[0]
            H, W = int(s.image_height), int(s.image_width)
            if colors_precomp.numel() and colors_precomp.shape[-1] != C_channels:
                raise RuntimeError(f"colors_precomp must have {C_channels} channels, got {tuple(colors_precomp.shape)}")

            m3 = _f32c(means3D, device)
            shc, col, opa = _f32c(sh, device), _f32c(colors_precomp, device), _f32c(opacities, device)
            sc, rot, cov = _f32c(scales, device), _f32c(rotations, device), _f32c(cov3Ds_precomp, device)
            bg = _pad_bg(s.bg, C_channels, device)
            vm, pm, cam = _f32c(s.viewmatrix, device), _f32c(s.projmatrix, device), _f32c(s.campos, device)

            # every pixel / every radius is written by the kernels (P == 0 excepted); only the atomically accumulated
            # blend weights need a zero fill
            alloc = torch.empty if P > 0 else torch.zeros
            color = alloc((C_channels, H, W), dtype=torch.float32, device=device)
            others = alloc((7, H, W), dtype=torch.float32, device=device)
            weight = torch.zeros((P, 1), dtype=torch.float32, device=device)
            radii = alloc((P,), dtype=torch.int32, device=device)
            scratch = _Scratch(device, (H, W, C_channels))

            args = _make_args(C_channels, s, bg, m3, shc, col, opa, sc, rot, cov, vm, pm, cam)
            budget, count_dev = None, None
            if _SYNC_FREE["enabled"] and P > 0:
                budget = _BUDGETS.setdefault((str(device), H, W, C_channels, P), _InstanceBudget())
                budget.poll()
                args.instance_capacity = budget.capacity()
                if args.instance_capacity:
                    count_dev = torch.empty(2, dtype=torch.int32, device=device)
                    args.instance_count_out = _lib.ptr(count_dev)
            out = _lib.RasterOut(_lib.ptr(color), _lib.ptr(others), _lib.ptr(weight), _lib.ptr(radii))
            stream = torch.cuda.current_stream(device).cuda_stream

            def run():
                with torch.cuda.device(device):
                    return _lib.check(L.envgs_raster_forward(C.byref(args), C.byref(out), scratch.cbs["geom"],
                                                             scratch.cbs["binning"], scratch.cbs["img"], None, stream),
                                      "envgs_raster_forward")

            if s.debug:
                cpu_args = cpu_deep_copy_tuple((s.bg, means3D, colors_precomp, opacities, scales, rotations,
                                                s.scale_modifier, cov3Ds_precomp, s.viewmatrix, s.projmatrix, s.tanfovx,
                                                s.tanfovy, s.image_height, s.image_width, sh, s.sh_degree, s.campos,
                                                s.prefiltered, s.debug))
                try:
                    num_rendered = run()
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                num_rendered = run()
            if budget is not None:
                if count_dev is None:
                    budget.max_seen = max(budget.max_seen, int(num_rendered))   # the synchronous first frame
                else:
                    with torch.cuda.device(device):
                        budget.watch(count_dev)

            ctx.raster_settings = s
            ctx.num_rendered = num_rendered
            ctx.bg = bg
            geom_buf, binning_buf, img_buf = scratch.get("geom"), scratch.get("binning"), scratch.get("img")
            scratch.close()
            ctx.save_for_backward(col, m3, sc, rot, cov, radii, shc, opa, geom_buf, binning_buf, img_buf)
            ctx.mark_non_differentiable(radii)
            return color, radii, others, weight

        @staticmethod
        def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_weight):
            L = _lib.lib()
            s = ctx.raster_settings
            col, m3, sc, rot, cov, radii, shc, opa, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
            device = m3.device
            P = m3.shape[0]
            H, W = int(s.image_height), int(s.image_width)
            M = shc.shape[1] if shc.numel() else 0

            g_color = (torch.zeros((C_channels, H, W), dtype=torch.float32, device=device) if grad_out_color is None
                       else grad_out_color.to(torch.float32).contiguous())
            g_others = (torch.zeros((7, H, W), dtype=torch.float32, device=device) if grad_depth is None
                        else grad_depth.to(torch.float32).contiguous())

            z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)
            zs = lambda saved, *shape: gradsink.zeros_or_sink(saved, shape, dtype=torch.float32, device=device)   # flat-buffer slice if registered
            dL_dmeans3D, dL_dmeans2D, dL_dcolors = zs(m3, P, 3), z(P, 3), z(P, C_channels)
            dL_dnormal, dL_dopacity, dL_dtransMat = z(P, 3), zs(opa, P, 1), z(P, 9)
            dL_dsh = zs(shc, P, M, 3) if M else z(P, M, 3)
            dL_dscales = zs(sc, P, 2) if sc.numel() else z(P, 2)
            dL_drotations = zs(rot, P, 4) if rot.numel() else z(P, 4)

            if P != 0:
                vm, pm, cam = _f32c(s.viewmatrix, device), _f32c(s.projmatrix, device), _f32c(s.campos, device)
                args = _make_args(C_channels, s, ctx.bg, m3, shc, col, opa, sc, rot, cov, vm, pm, cam)
                grads = _lib.RasterGrads(_lib.ptr(g_color), _lib.ptr(g_others), _lib.ptr(dL_dmeans2D),
                                         _lib.ptr(dL_dcolors), _lib.ptr(dL_dopacity), _lib.ptr(dL_dmeans3D),
                                         _lib.ptr(dL_dtransMat), _lib.ptr(dL_dsh), _lib.ptr(dL_dscales),
                                         _lib.ptr(dL_drotations), _lib.ptr(dL_dnormal))
                stream = torch.cuda.current_stream(device).cuda_stream

                def run():
                    with torch.cuda.device(device):
                        _lib.check(L.envgs_raster_backward(C.byref(args), _lib.ptr(radii), ctx.num_rendered,
                                                           _lib.ptr(geomBuffer), _lib.ptr(binningBuffer),
                                                           _lib.ptr(imgBuffer), C.byref(grads), stream),
                                   "envgs_raster_backward")

                if s.debug:
                    cpu_args = cpu_deep_copy_tuple((s.bg, m3, radii, col, sc, rot, s.scale_modifier, cov, s.viewmatrix,
                                                    s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color, grad_depth, shc,
                                                    s.sh_degree, s.campos, geomBuffer, ctx.num_rendered, binningBuffer,
                                                    imgBuffer, s.debug))
                    try:
                        run()
                    except Exception as ex:
                        torch.save(cpu_args, "snapshot_bw.dump")
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex
                else:
                    run()

            def like(g, ref):  # empty placeholder inputs receive empty grads
                return g if ref.numel() else None

            return (dL_dmeans3D, dL_dmeans2D, like(dL_dsh, shc), like(dL_dcolors, col), dL_dopacity, like(dL_dscales, sc),
                    like(dL_drotations, rot), like(dL_dtransMat, cov), None)

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings):
        return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, raster_settings)

    class GaussianRasterizer(nn.Module):
        NUM_CHANNELS = C_channels

        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def markVisible(self, positions):
            with torch.no_grad():
                s = self.raster_settings
                pos = positions.detach().to(torch.float32).contiguous()
                P = pos.shape[0]
                present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
                if P:
                    vm = s.viewmatrix.detach().to(torch.float32).contiguous()
                    pm = s.projmatrix.detach().to(torch.float32).contiguous()
                    with torch.cuda.device(pos.device):
                        _lib.check(_lib.lib().envgs_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                                                 present.data_ptr(),
                                                                 torch.cuda.current_stream(pos.device).cuda_stream),
                                   "envgs_mark_visible")
            return present

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            raster_settings = self.raster_settings
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                    ((scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
            empty = lambda: torch.empty(0, dtype=torch.float32, device=means3D.device)
            if shs is None: shs = empty()
            if colors_precomp is None: colors_precomp = empty()
            if scales is None: scales = empty()
            if rotations is None: rotations = empty()
            if cov3D_precomp is None: cov3D_precomp = empty()
            return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                       raster_settings)

    _RasterizeGaussians.__name__ = "_RasterizeGaussians"
    return rasterize_gaussians, _RasterizeGaussians, GaussianRasterizer
