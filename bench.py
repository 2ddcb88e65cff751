#!/usr/bin/env python
"""bench.py — EnvGS per-view rendering hot path on B200: rasterize base surfels -> reflect -> trace env surfels -> blend,
forward + backward (+ one gradient all-reduce and a fused Adam step when N > 1 / always).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a kernels (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the oracle (CPU restatement) on the host cores

Workload (BASELINE.json configs[2]): 1M base + 256K env surfels, SH degree 3, 1080p, 1-bounce reflection trace, full
forward + backward, synthetic scene, one view per rank per step (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (P_base, P_env, H, W)
    "config3": (1_000_000, 262_144, 1080, 1920),
    "config2": (500_000, 0, 1080, 1920),
    "config1": (10_000, 0, 256, 256),
    "tiny": (20_000, 8_192, 270, 480),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region through NVML (a few microseconds per query; an
    `nvidia-smi -lms` loop holds driver locks long enough to stall this launch-heavy step)."""

    def __init__(self, gpu_index: int, period_s: float = 0.05):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.reasons, self.max_clock = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_clock = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:  # no NVML: report it instead of inventing numbers
            self.err = f"nvml unavailable: {ex}"
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as ex:
                self.err = str(ex)
                return
            self._stop.wait(self.period)

    def stop(self) -> dict:
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_clock, "reasons": sorted(self.reasons),
               "samples": len(sm)}
        if self.err:
            out["note"] = self.err
        return out


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample(workload: str, n_tiles: int = 24, n_rays: int = 48, seed: int = 0) -> dict:
    """One bounded forward+backward sample of the workload through the CPU oracle (float32, host threads):
    `n_tiles` 16x16 tiles of the base rasterization (every base surfel is preprocessed, the tile lists are exact) +
    `n_rays` reflection rays through the brute-force env tracer.  Returns seconds per pixel of a full view."""
    from envgs_b200 import scene
    from oracle import raster_oracle as RO
    from oracle import trace_oracle as TO
    from tests import helpers as Hh
    Pb, Pe, H, W = WORKLOADS[workload]
    threads = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    s0 = 0.005 if H >= 1080 else 0.005 * 1080 / H
    base = scene.make_base_set(Pb, seed=0, s0=s0)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    feats = Hh.base_features(base, cam, 5)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    g = torch.Generator().manual_seed(seed)
    tiles = torch.randperm(gx * gy, generator=g)[:n_tiles].tolist()
    leaf = lambda t: t.clone().requires_grad_(True)
    t0 = time.perf_counter()
    # per-surfel stage over ALL surfels (this is what one view costs) ...
    pre = RO.preprocess(base.means, base.scales, base.rotations, cam.viewmatrix, cam.projmatrix, W, H)
    t_pre = time.perf_counter() - t0
    # ... then composite only the sampled tiles, from the surfels whose rectangle touches them
    rx0, ry0, rx1, ry1 = pre["rect"]
    touch = torch.zeros(Pb, dtype=torch.bool)
    for t in tiles:
        ty, tx = divmod(t, gx)
        touch |= pre["valid"] & (rx0 <= tx) & (tx < rx1) & (ry0 <= ty) & (ty < ry1)
    idx = torch.nonzero(touch)[:, 0]
    t0 = time.perf_counter()
    means, opac, scales, rots, colors = (leaf(base.means[idx]), leaf(base.opacities[idx]), leaf(base.scales[idx]),
                                         leaf(base.rotations[idx]), leaf(feats[idx]))
    out = RO.rasterize(means, opac, scales, rots, colors_precomp=colors, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                       campos=cam.campos, bg=torch.zeros(3), W=W, H=H, tiles=tiles)
    (out.color.abs().sum() + out.allmap.abs().sum()).backward()
    t_tiles = time.perf_counter() - t0
    sec_per_pix = t_pre / (H * W) * 2.0 + t_tiles / (n_tiles * 256)      # x2: the per-surfel stage also runs backward
    t_t, pix_t = 0.0, 0
    if Pe > 0:
        env = scene.make_env_set(Pe, seed=1, s0=s0)
        ro, rd = Hh.synthetic_rays(1, n_rays, seed=5)
        t0 = time.perf_counter()
        em, eo, es, er, esh = leaf(env.means), leaf(env.opacities), leaf(env.scales), leaf(env.rotations), leaf(env.shs)
        tr = TO.trace(ro, rd, em, eo, es, er, shs=esh, campos=cam.campos, bg=torch.zeros(3), max_hits=128, ray_chunk=16)
        tr.rgb.abs().sum().backward()
        t_t = time.perf_counter() - t0
        pix_t = n_rays
        sec_per_pix += t_t / pix_t
    return {"sec_per_pixel": sec_per_pix, "seconds": t_pre + t_tiles + t_t, "raster_s": t_pre + t_tiles, "trace_s": t_t,
            "threads": threads,
            "sample": f"{n_tiles} of {gx * gy} tiles composited fwd+bwd (per-surfel stage over all {Pb} base surfels) + {pix_t} rays "
                      f"brute-force traced fwd+bwd over {Pe} env surfels; fp32 torch oracle, {threads} threads; scaled to a full view"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Pb, Pe, H, W = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_sample(args.workload, n_tiles=4, n_rays=8)
    vals, secs = [], []
    sample = ""
    for k in range(args.steps):
        r = cpu_sample(args.workload, seed=k)
        vals.append(1.0 / r["sec_per_pixel"] / 1e6)
        secs.append(r["seconds"])
        sample = r["sample"]
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "render Mpix/s (fwd+bwd, rasterize+trace)", "value": v, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {Pb} base + {Pe} env surfels, {W}x{H}, fwd+bwd, bounded CPU sample per step"},
        "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": min(cores, 32), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    from envgs_b200 import _lib, scene
    from envgs_b200.dist import ViewParallelTrainer, init_process_group_from_env
    from envgs_b200.pipeline import EnvGSRenderer
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
    rank, local, world = init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    Pb, Pe, H, W = WORKLOADS[args.workload]
    s0 = 0.005 if H >= 1080 else 0.005 * 1080 / H

    base = scene.make_base_set(Pb, seed=0, s0=s0)
    env = scene.make_env_set(Pe, seed=1, s0=s0) if Pe > 0 else None
    renderer = EnvGSRenderer(base, env, device=dev)
    trainer = ViewParallelTrainer(renderer, lr=1e-6, average=True, optimizer=True)
    n_views = 138                                         # sedan-sized view list
    cams_cpu = scene.ring_cameras(n_views, H, W, seed=2)
    gt = torch.Generator().manual_seed(3)
    target_host = torch.rand(H, W, 3, generator=gt).pin_memory()
    e2e_warm = max(3, args.warmup)                        # the host-fed path has its own allocator / launch pattern: warm it too
    total_steps = args.warmup + 2 * args.steps + e2e_warm

    def view_of(step):                                    # rank r renders views r, r+world, ... of the step's batch
        return cams_cpu[(step * world + rank) % n_views]

    def cam_host_pack(cam):                               # the per-step host inputs of the e2e path
        return torch.cat([cam.viewmatrix.reshape(-1), cam.projmatrix.reshape(-1), cam.campos.reshape(-1), cam.K.reshape(-1),
                          cam.R.reshape(-1), cam.T.reshape(-1)]).pin_memory()

    def cam_from_pack(cam, pack_dev):
        v, p, c = pack_dev[0:16].view(4, 4), pack_dev[16:32].view(4, 4), pack_dev[32:35]
        K, R, T = pack_dev[35:44].view(3, 3), pack_dev[44:53].view(3, 3), pack_dev[53:56].view(3, 1)
        return scene.Camera(cam.H, cam.W, K, R, T, cam.tanfovx, cam.tanfovy, v, p, c, cam.znear, cam.zfar)

    packs = [cam_host_pack(view_of(s)) for s in range(total_steps)]
    target_dev = target_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(step):
        cam = view_of(step).to(dev) if step >= len(dev_cams) else dev_cams[step]
        return trainer.step(cam, target_dev)

    dev_cams = [view_of(s).to(dev) for s in range(total_steps)]

    def e2e_step(step):
        pack_dev = packs[step].to(dev, non_blocking=True)
        tgt = target_host.to(dev, non_blocking=True)
        cam = cam_from_pack(view_of(step), pack_dev)
        loss = trainer.step(cam, tgt)
        return loss.to("cpu", non_blocking=False)          # D2H read of the step's result

    # ---- warm-up ----
    for s in range(args.warmup):
        resident_step(s)
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.warmup, args.warmup + args.steps):
        resident_step(s)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end (host camera + target each step, loss read back) ----
    for s in range(args.warmup + args.steps, args.warmup + args.steps + e2e_warm):
        e2e_step(s)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for s in range(args.warmup + args.steps + e2e_warm, args.warmup + 2 * args.steps + e2e_warm):
        e2e_step(s)
    ev3.record()
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    pix_per_step = H * W * world
    value = pix_per_step / (ms_per_step * 1e-3) / 1e6
    e2e_value = pix_per_step / (ms_e2e / args.steps * 1e-3) / 1e6
    h2d = packs[0].numel() * 4 + target_host.numel() * 4
    d2h = 4

    # ---- per-kernel device times + roofline of render_fwd (rank 0, separate short run so events do not perturb the numbers) ----
    roofline, kernels = None, None
    if rank == 0 and not args.no_profile:
        _lib.profile_enable(True)
        nprof = 3
        stats = []
        for s in range(nprof):
            cam = dev_cams[s % len(dev_cams)]
            trainer.flat.zero_grad()
            out = renderer(cam, target_dev)
            renderer.loss(out, target_dev).backward()
            torch.cuda.synchronize()
            stats.append((int((out["radii"] > 0).sum()),))
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        kernels = {k: round(v[1] / v[0], 4) for k, v in rep.items()}
        # num_rendered of the last frame: read from the autograd context is awkward; recompute cheaply from the profile run
        N = renderer_last_num_rendered(renderer, dev_cams[(nprof - 1) % len(dev_cams)])
        P_v = stats[-1][0]
        C = 5
        alg_bytes = (64 + 4 * C) * N + (48 + 4 * C) * H * W + 4 * P_v      # SURVEY.md §8d render_fwd
        peak, peak_src = measured_peaks()
        t_ms = kernels.get("render_fwd")
        if t_ms:
            ach = alg_bytes / (t_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            try:                                    # DRAM bytes of this kernel from the committed ncu --set full capture
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    tj = json.load(f).get(args.workload)
                if tj and tj.get("kernel") == "render_fwd_kernel<5>":
                    traffic, traffic_src = int(tj["traffic"]), tj.get("source")
            except Exception:
                pass
            roofline = {"kernel": "render_fwd_kernel<5>", "bound": "hbm", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "algorithmic_bytes": alg_bytes,
                        "launch_ms": t_ms, "num_rendered": N, "visible_surfels": P_v,
                        "note": "render_fwd is fp32-issue bound, not HBM bound (SURVEY.md §8d); see DESIGN.md"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_sample(args.workload)
        cpu = {"value": round(1.0 / r["sec_per_pixel"] / 1e6, 6), "unit": "Mpix/s", "cores": r["threads"], "kind": "port",
               "sample": r["sample"], "seconds": round(r["seconds"], 2)}

    line = {
        "metric": "render Mpix/s (fwd+bwd, rasterize+trace)", "value": round(value, 3), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {Pb} base + {Pe} env surfels, SH deg 3, {W}x{H}, 1-bounce reflection trace, "
                               f"fwd+bwd + BVH rebuild + grad all-reduce + fused Adam, 1 view/rank/step",
                   "l2": "inputs_larger_than_L2 (surfel parameters 240+61 MB, maps 25-265 MB each vs 126 MB L2)",
                   "parallelism": f"view-parallel x{world}"},
        "e2e": {"value": round(e2e_value, 3), "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels_ms": kernels,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def renderer_last_num_rendered(renderer, cam) -> int:
    """Instance count of one frame (what render_fwd streams): run the base pass once more and read it from autograd."""
    with torch.enable_grad():
        out = renderer.rasterize_base(cam)
        fn = out["render"].grad_fn
        # walk back to the rasterizer node, which stores num_rendered on its ctx
        seen, stack = set(), [fn]
        while stack:
            f = stack.pop()
            if f is None or f in seen:
                continue
            seen.add(f)
            if hasattr(f, "num_rendered"):
                return int(f.num_rendered)
            stack.extend(n for n, _ in f.next_functions)
    return 0


if __name__ == "__main__":
    main()
