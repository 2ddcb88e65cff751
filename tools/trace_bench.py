"""Time the tracer alone on the config-3 reflection rays (development aid, GPU box)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene, _lib
from envgs_b200.pipeline import EnvGSRenderer
import bench
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
base = scene.make_base_set(Pb, seed=0, s0=0.005); env = scene.make_env_set(Pe, seed=1, s0=0.005)
r = EnvGSRenderer(base, env, device=dev)
cam = scene.ring_cameras(4, H, W, seed=2)[0].to(dev)
target = torch.rand(H, W, 3, device=dev)
def run(n):
    for it in range(n):
        for p in r.parameters(): p.grad = None
        out = r(cam, target); r.loss(out, target).backward()
    torch.cuda.synchronize()
run(3)
for thr in (os.environ.get("THRESHOLDS") or "28").split(","):
    os.environ["ENVGS_TRACE_REFILL"] = thr
    run(1)
    _lib.profile_enable(True)
    run(5)
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    print(os.environ.get("ENVGS_B200_LIB", "main"), "refill", thr, {k: round(v[1] / v[0], 3) for k, v in rep.items() if "trace" in k or "render" in k}, flush=True)
