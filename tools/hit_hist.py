"""Histogram of quad hits per reflection ray at config 3 (from the tracer's hit cache); -1 = handed to the general kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from envgs_b200.pipeline import EnvGSRenderer
import bench
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
r = EnvGSRenderer(scene.make_base_set(Pb, seed=0, s0=0.005), scene.make_env_set(Pe, seed=1, s0=0.005), device=dev)
cam = scene.ring_cameras(4, H, W, seed=2)[0].to(dev)
out = r(cam, torch.rand(H, W, 3, device=dev))
ctx = out["env"]["rgb"].grad_fn
R = H * W
counts = ctx.hit_cache[R * 16 * 8: R * 16 * 8 + 4 * R].view(torch.int32)
vals, n = torch.unique(counts, return_counts=True)
for v, c in zip(vals.tolist(), n.tolist()):
    print(f"{v:3d}: {c:8d}  {100.0 * c / R:6.2f} %")
