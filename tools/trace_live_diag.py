"""Development aid (GPU box): three-way comparison at the live-test size (256 K env surfels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from tests import helpers as Hh
H, W = 96, 128
P = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
env = scene.make_env_set(P, seed=1, s0=0.005)
cam = scene.ring_cameras(1, H, W, seed=2)[0]
ro, rd = Hh.synthetic_rays(H, W, seed=5)
kw = dict(use_sh=True)
ref = Hh.run_reference_trace(env, cam, ro, rd, **kw)
cu = Hh.run_cuda_trace(env, cam, ro, rd, **kw)
orc = Hh.run_oracle_trace(env, cam, ro, rd, device="cuda", max_hits=256, ray_chunk=128, **kw)
for a, b, n in ((cu, ref, "cuda-ref"), (orc, ref, "oracle-ref"), (cu, orc, "cuda-oracle")):
    for k in ("rgb", "acc", "dpt"):
        d = (a[k] - b[k]).abs().amax(-1)
        print(n, k, f"max {float(d.max()):.3e} mean {float(d.mean()):.3e} frac>1e-4 {float((d > 1e-4).float().mean()):.3e}")
d = (cu["rgb"] - ref["rgb"]).abs().amax(-1)
print("border frac", float(orc["border"].float().mean()), "bad&border", int(((d > 1e-4) & orc["border"]).sum()), "bad", int((d > 1e-4).sum()))
idx = d.flatten().topk(5).indices
for i in idx:
    r, c = int(i) // W, int(i) % W
    print((r, c), "cu", cu["rgb"][r, c].tolist(), cu["acc"][r, c].item(), "ref", ref["rgb"][r, c].tolist(), ref["acc"][r, c].item(), "orc", orc["rgb"][r, c].tolist(), orc["acc"][r, c].item())
dw = (cu["wet"] - ref["wet"]).abs().flatten()
print("wet rel", Hh.rel_l2(cu["wet"], ref["wet"]), "top", [(int(i), float(v), float(ref["wet"].flatten()[i])) for v, i in zip(*dw.topk(5))])
dwo = (orc["wet"] - ref["wet"]).abs().flatten()
print("wet oracle rel", Hh.rel_l2(orc["wet"], ref["wet"]))
