"""Per-step wall time + NVML SM/memory clocks + power right after start-up: what is the slow ramp of the first ~50 steps?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, pynvml
from envgs_b200 import scene
from envgs_b200.pipeline import EnvGSRenderer
from envgs_b200.dist import ViewParallelTrainer
import bench
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
def clocks():
    return (pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM),
            pynvml.nvmlDeviceGetPowerUsage(hnd) // 1000, pynvml.nvmlDeviceGetPerformanceState(hnd))
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
print("idle", clocks())
r = EnvGSRenderer(scene.make_base_set(Pb, seed=0, s0=0.005), scene.make_env_set(Pe, seed=1, s0=0.005), device=dev)
tr = ViewParallelTrainer(r, lr=1e-6)
cams = [c.to(dev) for c in scene.ring_cameras(16, H, W, seed=2)]
target = torch.rand(H, W, 3, device=dev)
torch.cuda.synchronize()
print("after setup", clocks(), "reserved MB", torch.cuda.memory_reserved() >> 20)
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "preheat":      # busy the GPU for ~1.5 s with something unrelated first
    x = torch.randn(8192, 8192, device=dev)
    t = time.perf_counter()
    while time.perf_counter() - t < 1.5:
        y = x @ x
    torch.cuda.synchronize()
    print("after preheat", clocks())
if mode == "nosync":       # the bench pattern: no host sync between steps; per-step GPU intervals from events
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(71)]
    cpu = []
    evs[0].record()
    for s in range(70):
        t = time.perf_counter()
        tr.step(cams[s % 16], target)
        evs[s + 1].record()
        cpu.append((time.perf_counter() - t) * 1e3)
    torch.cuda.synchronize()
    for s in range(70):
        if s < 12 or s % 5 == 0:
            print(f"step {s:3d} gpu-interval {evs[s].elapsed_time(evs[s + 1]):7.2f} ms  cpu {cpu[s]:7.2f} ms")
    import gc
    a0 = torch.cuda.memory_allocated() >> 20; n = gc.collect(); print("gc.collect() ->", n, "objects; allocated", a0, "->", torch.cuda.memory_allocated() >> 20, "MB")
    st = torch.cuda.memory_stats(); print("cudaMallocs", st["num_device_alloc"], "reserved", torch.cuda.memory_reserved() >> 20, "allocated", torch.cuda.memory_allocated() >> 20, "peak", torch.cuda.max_memory_allocated() >> 20)
    segs = sorted(torch.cuda.memory_snapshot(), key=lambda x: -x["total_size"])[:12]
    for sg in segs: print("  segment", sg["total_size"] >> 20, "MB active", sg["active_size"] >> 20, "blocks", len(sg["blocks"]))
    sys.exit(0)
stats0 = torch.cuda.memory_stats()
for s in range(70):
    torch.cuda.synchronize(); t = time.perf_counter()
    tr.step(cams[s % 16], target)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    st = torch.cuda.memory_stats()
    if s < 12 or s % 5 == 0:
        print(f"step {s:3d} {dt:7.2f} ms  clocks {clocks()}  reserved {torch.cuda.memory_reserved() >> 20} MB  cudaMallocs {st['num_device_alloc']}  frees {st['num_device_free']}")
