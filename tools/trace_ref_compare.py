"""Development aid (GPU box): three-way comparison reference tracer (oracle/_ref/ref_trace) / fp64 oracle / CUDA kernels
on the golden trace cases."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import make_golden as MG
from tests import helpers as Hh

rep = {}
for name in MG.TRACE_CASES:
    env, cam, ro, rd, up, kw = MG.trace_case_inputs(name)
    ref = Hh.run_reference_trace(env, cam, ro, rd, grads=up, **kw)
    cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up, **kw)
    okw = dict(kw)
    if "dense" in name:
        okw.update(max_hits=1024, ray_chunk=256)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, grads=up, **okw)
    r = {}
    for k in ref:
        if not torch.is_tensor(ref[k]): continue
        scale = max(1.0, float(ref[k].abs().max()))
        for tag, other in (("cuda", cu), ("oracle", orc)):
            if k not in other: continue
            d = (other[k].double() - ref[k].double()).abs()
            r[f"{k}:{tag}"] = dict(max=float(d.max()), frac_gt_1e4=float((d > 1e-4 * scale).double().mean()),
                                   rel_l2=Hh.rel_l2(other[k], ref[k]))
    rep[name] = r
    print("==", name)
    for k, v in r.items():
        print(f"  {k:20s} max {v['max']:.3e} frac>1e-4 {v['frac_gt_1e4']:.3e} relL2 {v['rel_l2']:.3e}")
with open(os.path.join(ROOT, "gpurun_out", "trace_ref_compare.json"), "w") as f:
    json.dump(rep, f, indent=1)
