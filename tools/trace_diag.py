"""Development aid (GPU box): locate the ray / surfel behind the largest CUDA-vs-reference difference of a trace golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import make_golden as MG
from oracle.raster_oracle import quat_to_rotmat
from tests import helpers as Hh
name = sys.argv[1] if len(sys.argv) > 1 else "trace_ref_two_segments"
gold = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"))
env, cam, ro, rd, up, kw = MG.trace_case_inputs(name)
cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up, **kw)
d = (cu["rgb"] - gold["rgb"]).abs().amax(-1)
H, W = d.shape
top = d.flatten().topk(3)
dw = (cu["wet"] - gold["wet"]).abs().flatten()
print("wet top", [(int(i), float(v)) for v, i in zip(*dw.topk(4))])
for v, idx in zip(top.values, top.indices):
    i, j = int(idx) // W, int(idx) % W
    print(f"pixel ({i},{j}) diff {float(v):.3e}\n  cuda mid {cu['mid'][i, j].numpy().round(5)}\n  gold mid {gold['mid'][i, j].numpy().round(5)}")
    for seg in range(kw.get("max_trace_depth", 0) + 1):
        o = gold["mid"][i, j, 16 * seg:16 * seg + 3].double(); dd = gold["mid"][i, j, 16 * seg + 3:16 * seg + 6].double()
        if float(dd.norm()) == 0: continue
        means, sc, rot, op = env.means.double(), env.scales.double(), env.rotations.double(), env.opacities.double()[:, 0]
        R = quat_to_rotmat(rot); tu, tv = R[:, :, 0], R[:, :, 1]; m = torch.linalg.cross(tu, tv)
        t = ((means - o) * m).sum(-1) / (m * dd).sum(-1)
        x = o + t[:, None] * dd
        u = ((x - means) * tu).sum(-1) / sc[:, 0]; vv = ((x - means) * tv).sum(-1) / sc[:, 1]
        alpha = op * torch.exp(-0.5 * (u * u + vv * vv))
        hit = (t > 0) & (u.abs() <= 3.001) & (vv.abs() <= 3.001)
        ids = hit.nonzero()[:, 0]
        order = t[ids].argsort()
        print(f"  segment {seg}: {len(ids)} hits")
        T = 1.0
        for k in ids[order][:40]:
            k = int(k)
            a = min(0.99, float(alpha[k]))
            flag = "" if a >= 1 / 255 else " (alpha<1/255)"
            if float(t[k]) < 0.06: flag += " (t<0.06)"
            print(f"    id {k:5d} t {float(t[k]):.7f} u {float(u[k]):+.6f} v {float(vv[k]):+.6f} alpha {a:.7f} T {T:.6f} wetdiff {float(dw[k]):.2e}{flag}")
            if a >= 1 / 255 and float(t[k]) >= 0.06: T *= (1 - a)
