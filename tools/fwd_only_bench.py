"""Forward-only timing of the tracer kernels at config 3 (experiments; results of EXP_* builds are not valid renders)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene, _lib
from envgs_b200.pipeline import EnvGSRenderer
import bench
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
r = EnvGSRenderer(scene.make_base_set(Pb, seed=0, s0=0.005), scene.make_env_set(Pe, seed=1, s0=0.005), device=dev)
cam = scene.ring_cameras(4, H, W, seed=2)[0].to(dev)
target = torch.rand(H, W, 3, device=dev)
with torch.no_grad():
    for it in range(3): r(cam, target)
    torch.cuda.synchronize(); _lib.profile_enable(True)
    for it in range(5): r(cam, target)
    torch.cuda.synchronize()
rep = _lib.profile_report()
print(os.environ.get("ENVGS_B200_LIB", "main"), {k: round(v[1] / v[0], 3) for k, v in rep.items() if "trace" in k})
