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
import gc
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

TRACE_DEPTH = {"config5": 1}      # trace segments beyond the first (BASELINE config 5: 2-bounce)
WORKLOADS = {
    # name: (P_base, P_env, H, W)
    "config3": (1_000_000, 262_144, 1080, 1920),
    "config5": (4_000_000, 1_048_576, 2160, 3840),
    "config2": (500_000, 0, 1080, 1920),
    "config1": (10_000, 0, 256, 256),
    "tiny": (20_000, 8_192, 270, 480),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: do not start the environment group's all-reduce from the gradient hook")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: all-reduce + full-size Adam instead of reduce-scatter / sharded Adam / all-gather")
    ap.add_argument("--no-factored-sh", action="store_true", help="N > 1: exchange the base SH gradient itself instead of its rank-one factors")
    ap.add_argument("--raster-sync-free", action="store_true",
                    help="rasterizer without the reference's blocking read of the (tile, surfel) instance count (capacity mode after the "
                         "first frame of a shape; measured: no gain on this GPU-bound step, see DESIGN.md section 8b)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the in-process timing of the reference GPU operators")
    ap.add_argument("--reference-device", default="auto", choices=["auto", "cpu"],
                    help="--impl reference: 'auto' runs the real reference on the GPU when it can (else the CPU port); 'cpu' forces the CPU port")
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
    s0 = 0.005 * 1080 / H          # mean 3-sigma screen radius ~6 px at every resolution (s0 = 6 z / (3 f), f = 0.78 W)
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


def reference_gpu_renderer(Pb, Pe, H, W, dev):
    """The per-view path on the UNMODIFIED reference operators (oracle/ref_ops.py: the reference rasterizer extension built
    for sm_100a + its tracer device programs under the driver's real OptiX) with the reference's own PyTorch glue
    (EnvGSRenderer fused_* = False mirrors envgs_sampler.py / gaussian2d_utils.py line by line).  None of this repository's
    kernels run on this path."""
    from envgs_b200 import scene
    from envgs_b200.pipeline import EnvGSRenderer
    from oracle import ref_ops
    s0 = 0.005 * 1080 / H          # mean 3-sigma screen radius ~6 px at every resolution (s0 = 6 z / (3 f), f = 0.78 W)
    base = scene.make_base_set(Pb, seed=0, s0=s0)
    env = scene.make_env_set(Pe, seed=1, s0=s0) if Pe > 0 else None
    return EnvGSRenderer(base, env, device=dev, fused_features=False, fused_glue=False, rasterizer_cls=ref_ops.RefGaussianRasterizer,
                         tracer=ref_ops.RefSurfelTracer() if Pe > 0 else None)


def run_reference_arm_gpu(args) -> bool:
    """`--impl reference` when the real reference can run on this box's GPU: full steps of the same workload (render +
    backward + torch Adam, as the reference's trainer does), timed like the main arm.  Returns False if unavailable."""
    if args.reference_device == "cpu" or not torch.cuda.is_available():
        return False
    from oracle import ref_ops
    ok, why = ref_ops.available()
    if not ok:
        print(f"[bench] real reference not runnable on the GPU here ({why}); falling back to the CPU port", file=sys.stderr)
        return False
    from envgs_b200 import scene
    Pb, Pe, H, W = WORKLOADS[args.workload]
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    renderer = reference_gpu_renderer(Pb, Pe, H, W, dev)
    opt = torch.optim.Adam(renderer.parameters(), lr=1e-6, eps=1e-15)
    cams = [c.to(dev) for c in scene.ring_cameras(138, H, W, seed=2)]
    target = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(3)).to(dev)

    def step(k):
        opt.zero_grad(set_to_none=True)
        out = renderer(cams[k % len(cams)], target)
        loss = renderer.loss(out, target)
        loss.backward()
        opt.step()
        return loss

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()                                          # same host-side policy as the main arm: no cyclic GC pass inside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    gc.enable()
    ms = e0.elapsed_time(e1) / args.steps
    v = H * W / (ms * 1e-3) / 1e6
    from oracle import ref_trace as RT
    sample = (f"full steps of the same workload on the UNMODIFIED reference operators on this box's GPU 0 (the reference has no CPU "
              f"path, SURVEY.md §8d): diff-surfel-rasterization-wet-ch05 CUDA extension built for sm_100a + diff-surfel-tracing device "
              f"programs under the driver's OptiX runtime ({RT.optix_info().get('nvoptix')}, PTX {RT.optix_info().get('ptx')}), reference "
              f"PyTorch glue, torch.optim.Adam; 1 view/step")
    line = {
        "impl": "reference", "metric": "render Mpix/s (fwd+bwd, rasterize+trace)", "value": round(v, 4), "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {Pb} base + {Pe} env surfels, SH deg 3, {W}x{H}, 1-bounce reflection trace, fwd+bwd + "
                               f"OptiX GAS rebuild + Adam, 1 view/step on 1 GPU (the reference has no multi-GPU path)",
                   "reference_device": "gpu"},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mpix/s", "cores": 0, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 4), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return True


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if run_reference_arm_gpu(args):
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
        "config": {"workload": f"{args.workload}: {Pb} base + {Pe} env surfels, {W}x{H}, fwd+bwd, bounded CPU sample per step",
                   "reference_device": "cpu"},
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

    from envgs_b200 import _lib, rasterizer, scene
    from envgs_b200.dist import ViewParallelTrainer, init_process_group_from_env
    from envgs_b200.pipeline import EnvGSRenderer
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
    rank, local, world = init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    Pb, Pe, H, W = WORKLOADS[args.workload]
    rasterizer.set_sync_free(args.raster_sync_free)
    s0 = 0.005 * 1080 / H          # mean 3-sigma screen radius ~6 px at every resolution (s0 = 6 z / (3 f), f = 0.78 W)

    base = scene.make_base_set(Pb, seed=0, s0=s0)
    env = scene.make_env_set(Pe, seed=1, s0=s0) if Pe > 0 else None
    if TRACE_DEPTH.get(args.workload, 0) and env is not None:
        env.specular.fill_(0.5)                           # every second-segment ray continues unless its first segment saw nothing
    renderer = EnvGSRenderer(base, env, device=dev, max_trace_depth=TRACE_DEPTH.get(args.workload, 0))
    trainer = ViewParallelTrainer(renderer, lr=1e-6, average=True, optimizer=True, overlap=not args.no_overlap,
                                  shard_optimizer=not args.no_shard, factored_sh=not args.no_factored_sh)
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

    class HostFeed:
        """Per-step host inputs (camera block + 25 MB target) copied from pinned memory on a side stream, one step ahead: step
        k's copy is issued while step k-1 computes and step k only waits on its own copy's event.  Every step's inputs still
        cross PCIe inside the timed region; they just no longer sit on the critical path."""

        def __init__(self):
            self.stream = torch.cuda.Stream(dev)
            self.slots = {}

        def issue(self, step):
            if step in self.slots or step >= total_steps:
                return
            with torch.cuda.stream(self.stream):
                pack_dev = packs[step].to(dev, non_blocking=True)
                tgt = target_host.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.slots[step] = (pack_dev, tgt, ev)

        def get(self, step):
            self.issue(step)
            pack_dev, tgt, ev = self.slots.pop(step)
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            pack_dev.record_stream(cur)
            tgt.record_stream(cur)
            return pack_dev, tgt

    feed = HostFeed()
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def e2e_step(step):
        pack_dev, tgt = feed.get(step)
        feed.issue(step + 1)
        cam = cam_from_pack(view_of(step), pack_dev)
        trainer.step(cam, tgt, loss_out=loss_host)
        trainer.loss_ready.synchronize()                   # D2H read of the step's result: the copy is queued right behind the loss kernel,
        return float(loss_host)                            # so the host waits for the forward pass only and goes on queueing the next step

    # ---- warm-up ----
    for s in range(args.warmup):
        resident_step(s)
    barrier()
    # Manual GC (as large training loops do): one collection here, none inside the timed regions — a generation-2 pass of the cyclic
    # collector is a multi-millisecond host stall that lands on a random step.
    gc.collect()
    gc.disable()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_events = []                                      # ENVGS_BENCH_STEP_TIMES=1: per-step device times (view-cost spread)
    barrier()
    ev0.record()
    for s in range(args.warmup, args.warmup + args.steps):
        resident_step(s)
        if os.environ.get("ENVGS_BENCH_STEP_TIMES") == "1":
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            step_events.append(e)
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
    gc.enable()

    if os.environ.get("ENVGS_BENCH_TORCH_PROFILE"):      # diagnostic: per-kernel device time of 6 steps on rank 0 (after the timed regions)
        from torch.profiler import ProfilerActivity, profile
        barrier()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for s in range(6):
                resident_step(args.warmup + s)
            barrier()
        if rank == 0:
            with open(os.environ["ENVGS_BENCH_TORCH_PROFILE"], "w") as f:
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=90))
            if os.environ.get("ENVGS_BENCH_TORCH_TRACE"):
                prof.export_chrome_trace(os.environ["ENVGS_BENCH_TORCH_TRACE"])

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

    # ---- per-kernel device times + rooflines (rank 0, separate short run so the events do not perturb the timed region) ----
    roofline, rooflines, kernels, ref_gpu = None, None, None, None
    if rank == 0 and not args.no_profile:
        import ctypes
        have_env = Pe > 0
        _lib.profile_enable(True)
        nprof = 3
        P_v, P_hit = 0, 0
        for s in range(nprof):
            cam = dev_cams[s % len(dev_cams)]
            trainer.flat.zero_grad()
            out = renderer(cam, target_dev)
            renderer.loss(out, target_dev).backward()
            torch.cuda.synchronize()
            P_v = int((out["radii"] > 0).sum())
            if have_env:
                P_hit = int((out["env"]["wet"] > 0).sum())
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        kernels = {k: round(v[1] / v[0], 4) for k, v in rep.items()}
        Hh = 0
        if have_env:                                # one extra forward with the (slower) counting traversal, outside the timed profile
            _lib.lib().envgs_debug_trace_stats(renderer.tracer.optix_context.handle, 1, None)
            with torch.no_grad():
                renderer(dev_cams[(nprof - 1) % len(dev_cams)], target_dev)
            torch.cuda.synchronize()
            buf = (ctypes.c_uint64 * 8)()
            _lib.lib().envgs_debug_trace_stats(renderer.tracer.optix_context.handle, 0, buf)
            Hh = int(buf[3])                        # quad hits found by the traversal of one view (>= the composited hits)
        N = renderer_last_num_rendered(renderer, dev_cams[(nprof - 1) % len(dev_cams)])
        C, S, R = 5, 16, H * W
        peak, peak_src = measured_peaks()
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic_db = json.load(f).get(args.workload, {})
        except Exception:
            traffic_db = {}
        # SURVEY.md §8d algorithmic bytes per view; "kernels" lists the profile entries whose time the bytes are set against
        groups = {
            "render_fwd": dict(kernel="render_fwd_kernel<5>", kernels=["render_fwd"], bytes=(64 + 4 * C) * N + (48 + 4 * C) * H * W + 4 * P_v),
            "render_bwd": dict(kernel="render_bwd_kernel<5>", kernels=["render_bwd"],
                               bytes=(64 + 4 * C) * N + (48 + 4 * C) * H * W + (60 + 4 * C) * P_v),
        }
        if have_env:
            groups["trace_fwd"] = dict(kernel="traverse_kernel + shade_fwd_kernel (+ overflow fallback)",
                                       kernels=["trace_traverse", "trace_class_lists", "trace_shade_fwd", "trace_fwd_fallback"],
                                       bytes=24 * R + (52 + 64 * 1) * R + (44 + 12 * S) * Hh)
            groups["trace_bwd"] = dict(kernel="shade_bwd_kernel (+ overflow fallback)", kernels=["trace_shade_bwd", "trace_bwd_fallback"],
                                       bytes=(24 + 52 + 64 * 1 + 44 + 24) * R + (40 + 12 * S) * Hh + (40 + 12 * S) * P_hit)
            # the dominant single kernel on its own (24 B of ray in, one 128-byte hit line out per ray)
            groups["traverse"] = dict(kernel="traverse_kernel (alone; part of trace_fwd)", kernels=["trace_traverse"], bytes=(24 + 128) * R)
        rooflines = []
        for name, gdef in groups.items():
            t_ms = sum(kernels.get(k, 0.0) for k in gdef["kernels"])
            if t_ms <= 0:
                continue
            ach = gdef["bytes"] / (t_ms * 1e-3) / 1e9
            tj = traffic_db.get(name, {})
            rooflines.append({"name": name, "kernel": gdef["kernel"], "bound": "hbm", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s",
                              "frac": round(ach / peak, 4), "traffic": tj.get("traffic"), "traffic_source": tj.get("source"),
                              "peak_source": peak_src, "algorithmic_bytes": int(gdef["bytes"]), "launch_ms": round(t_ms, 4)})
        if rooflines:
            # the dominant stage of the step in SURVEY.md section 8(d)'s decomposition (trace_fwd = traversal + shading: its
            # per-unit byte figure covers both kernels); the traversal kernel alone is the `traverse` entry of `rooflines`
            roofline = dict(max(rooflines, key=lambda r: r["launch_ms"]))
            roofline.update(num_rendered=N, visible_surfels=P_v, rays=R, composited_hits=Hh, env_surfels_hit=P_hit,
                            note="every group is fp32-issue / latency bound, not HBM bound (SURVEY.md §8d; DESIGN.md §6): measured DRAM "
                                 "traffic is below the algorithmic bytes except for render_bwd (1.55x: read-modify-write of the gradient "
                                 "rows its REDs touch), at 4 % of the HBM peak either way")

        # ---- the UNMODIFIED reference operators on the same inputs, same process (oracle/ref_ops.py) ----
        if world == 1 and not args.no_ref_gpu:
            try:
                ref_gpu = time_reference_gpu_ops(renderer, dev_cams[0], target_dev, kernels)
            except Exception as ex:   # the reference build is a checker, never a dependency of the bench line
                ref_gpu = {"unavailable": f"{type(ex).__name__}: {ex}"}

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
        "config": {"workload": f"{args.workload}: {Pb} base + {Pe} env surfels, SH deg 3, {W}x{H}, {1 + TRACE_DEPTH.get(args.workload, 0)}-segment reflection trace, "
                               f"fwd+bwd + BVH rebuild + grad all-reduce + fused Adam, 1 view/rank/step",
                   "l2": "inputs_larger_than_L2 (surfel parameters 240+61 MB, maps 25-265 MB each vs 126 MB L2)",
                   "parallelism": f"view-parallel x{world}",
                   "raster_host_sync": not args.raster_sync_free, "host_gc": "one collection before the timed regions, cyclic GC off inside (both arms)"},
        "e2e": {"value": round(e2e_value, 3), "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "rooflines": rooflines,
        "ref_gpu": ref_gpu,
        "kernels_ms": kernels,
        "cpu_baseline": cpu,
    }
    if step_events:
        marks = [ev0] + step_events
        line["step_ms"] = [round(marks[i].elapsed_time(marks[i + 1]), 4) for i in range(len(step_events))]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


RASTER_KERNELS = ("preprocess_fwd", "scan", "duplicate", "sort", "tile_ranges", "gather_sorted", "render_fwd", "render_bwd", "preprocess_bwd")
TRACE_KERNELS = ("glue_quad_vertices", "bvh_build", "trace_traverse", "trace_class_lists", "trace_shade_fwd", "trace_fwd_fallback",
                 "trace_shade_bwd", "trace_bwd_fallback")


def time_reference_gpu_ops(renderer, cam, target, kernels_ms) -> dict:
    """Time the reference's two operators (forward + backward, its own syncs included: that is their real cost) on exactly
    the tensors this repository's operators see in one view: same base surfels / features / camera for the rasterizer, the
    same reflection rays and env surfels for the tracer (GAS build included: EnvGS rebuilds it every iteration)."""
    from envgs_b200 import glue, scene
    from envgs_b200.rasterizer import GaussianRasterizationSettings
    from envgs_b200.tracer import SurfelTracingSettings
    from envgs_b200.features import sh_features
    from oracle import ref_ops
    from oracle import ref_trace as RT
    dev = target.device
    out = {}

    def med(fn, warm=2, it=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(it):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    r = renderer
    with torch.no_grad():
        colors = sh_features(r.means, r.shs, torch.cat([r.specular, r.roughness], dim=-1), cam.campos, 3)
        image, radii, allmap, weight, _ = r._raster(cam)
        ref_o, ref_d = glue.reflect_rays(allmap, cam)
    settings = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, r.bg, 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.campos,
                                             False, False)
    g_img, g_all = torch.randn_like(image) * 1e-3, torch.randn_like(allmap) * 1e-3
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    st = {}

    def rast_fwd():
        st["leafs"] = [leaf(t) for t in (r.means, r.opacities, r.scales, r.rotations, colors)]
        m, o, sc, ro, co = st["leafs"]
        st["out"] = ref_ops.RefGaussianRasterizer(settings)(means3D=m, means2D=torch.zeros_like(m, requires_grad=True), opacities=o,
                                                            colors_precomp=co, scales=sc, rotations=ro)

    def rast_fwd_bwd():
        rast_fwd()
        torch.autograd.backward([st["out"][0], st["out"][2]], [g_img, g_all])

    try:
        f = med(rast_fwd)
        fb = med(rast_fwd_bwd)
        ours = sum(kernels_ms.get(k, 0.0) for k in RASTER_KERNELS)
        out["rasterizer"] = {"ref_fwd_ms": round(f, 3), "ref_fwd_bwd_ms": round(fb, 3), "ours_fwd_bwd_kernels_ms": round(ours, 3),
                             "speedup": round(fb / ours, 2) if ours else None,
                             "what": "diff-surfel-rasterization-wet-ch05 (unmodified, built for sm_100a) vs this repo's kernels "
                                     "(preprocess..render_bwd, CUDA-event device time), same surfels / features / camera"}
    except Exception as ex:   # noqa: BLE001
        out["rasterizer"] = {"unavailable": f"{type(ex).__name__}: {ex}"}

    if r.has_env:
        if RT.optix_available():
            ts = SurfelTracingSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, r.env_bg, 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.campos,
                                       False, False, 0, 0.0)
            tracer = ref_ops.RefSurfelTracer()
            g_rgb = torch.randn(cam.H, cam.W, 3, device=dev) * 1e-3

            def trace_fwd_bwd():
                with torch.no_grad():
                    v, fidx = scene.quad_vertices(r.env_means, r.env_scales, r.env_rotations)     # get_disks, optix_utils.py:39-69
                    tracer.build_acceleration_structure(v, fidx, rebuild=True)
                lf = [leaf(t) for t in (r.env_means, r.env_opacities, r.env_scales, r.env_rotations, r.env_shs)]
                o_, d_ = leaf(ref_o), leaf(ref_d)
                res = tracer(o_, d_, v, lf[0], torch.zeros_like(lf[0], requires_grad=True), shs=lf[4], opacities=lf[1], scales=lf[2],
                             rotations=lf[3], tracer_settings=ts, start_from_first=False)
                res[0].backward(g_rgb)

            fb = med(trace_fwd_bwd)
            ours = sum(kernels_ms.get(k, 0.0) for k in TRACE_KERNELS)
            out["tracer"] = {"ref_build_fwd_bwd_ms": round(fb, 3), "ours_build_fwd_bwd_kernels_ms": round(ours, 3),
                             "speedup": round(fb / ours, 2) if ours else None,
                             "what": f"diff-surfel-tracing device programs (unmodified PTX) under the driver's OptiX runtime "
                                     f"({RT.optix_info().get('nvoptix')}) vs this repo's LBVH tracer, same reflection rays / env surfels"}
        else:
            out["tracer"] = {"unavailable": str(RT.optix_info().get("err"))}
    if "speedup" in out.get("rasterizer", {}) and "speedup" in out.get("tracer", {}):
        ref_total = out["rasterizer"]["ref_fwd_bwd_ms"] + out["tracer"]["ref_build_fwd_bwd_ms"]
        our_total = out["rasterizer"]["ours_fwd_bwd_kernels_ms"] + out["tracer"]["ours_build_fwd_bwd_kernels_ms"]
        out["operators_total"] = {"ref_ms": round(ref_total, 3), "ours_ms": round(our_total, 3), "speedup": round(ref_total / our_total, 2)}
    return out


def renderer_last_num_rendered(renderer, cam) -> int:
    """Instance count of one frame (what render_fwd streams): run the base pass once more and read it from autograd.  The
    synchronous mode is forced for this call: in the sync-free mode the context holds the capacity, not the count."""
    from envgs_b200 import rasterizer
    was = dict(rasterizer._SYNC_FREE)
    rasterizer.set_sync_free(False)
    try:
        return _last_num_rendered(renderer, cam)
    finally:
        rasterizer.set_sync_free(was["enabled"], was["margin"])


def _last_num_rendered(renderer, cam) -> int:
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
