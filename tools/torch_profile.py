"""torch.profiler breakdown of one config-3 training step (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from envgs_b200 import scene
from envgs_b200.pipeline import EnvGSRenderer
from envgs_b200.dist import ViewParallelTrainer
import bench
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
r = EnvGSRenderer(scene.make_base_set(Pb, seed=0, s0=0.005), scene.make_env_set(Pe, seed=1, s0=0.005), device=dev)
tr = ViewParallelTrainer(r, lr=1e-6)
cams = [c.to(dev) for c in scene.ring_cameras(4, H, W, seed=2)]
target = torch.rand(H, W, 3, device=dev)
for s in range(3): tr.step(cams[s % 4], target)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(3): tr.step(cams[s % 4], target)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
