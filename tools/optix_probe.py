"""Development aid (GPU box): run the reference tracer's device programs under the REAL OptiX runtime
(oracle/ref_trace_optix_driver.cpp), compare with the emulation goldens, and time BASELINE config 3's tracer half."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from oracle import ref_trace as RT
from oracle import make_golden as MG
from tests import helpers as Hh
rep = {"available": RT.optix_available(), "info": RT.optix_info()}
print(rep, flush=True)
if not rep["available"]:
    sys.exit(0)
for name in MG.TRACE_CASES:
    gold = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"))
    env, cam, ro, rd, up, kw = MG.trace_case_inputs(name)
    out = Hh.run_reference_trace(env, cam, ro, rd, grads=up, backend="optix", **kw)
    r = {}
    for k in gold:
        if torch.is_tensor(gold[k]) and k in out and k != "mid":
            d = (out[k].double() - gold[k].double()).abs()
            r[k] = dict(max=float(d.max()), rel_l2=Hh.rel_l2(out[k], gold[k]), n_gt_1e4=int((d > 1e-4 * max(1.0, float(gold[k].abs().max()))).sum()))
    rep[name] = r
    print(name, {k: (f"{v['max']:.2e}", f"{v['rel_l2']:.2e}", v["n_gt_1e4"]) for k, v in r.items()}, flush=True)
# timing at BASELINE config 3's tracer half: 1080p reflection-like rays, 256 K env surfels, fwd + bwd + GAS build
if "--time" in sys.argv:
    H, W = 1080, 1920
    dev = torch.device("cuda")
    env = scene.make_env_set(262144, seed=1, s0=0.005).to(dev)
    cam = scene.ring_cameras(1, H, W, seed=2)[0].to(dev)
    ro, rd = Hh.synthetic_rays(H, W, seed=5)
    v, f = scene.quad_vertices(env.means, env.scales, env.rotations)
    kw = dict(ray_o=ro.to(dev), ray_d=rd.to(dev), vertices=v, triangles=f, bg=torch.zeros(3, device=dev), means3D=env.means, shs=env.shs, degree=3,
              colors_precomp=None, others_precomp=None, opacities=env.opacities, scales=env.scales, scale_modifier=1.0, rotations=env.rotations,
              viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos, start_from_first=False, max_trace_depth=0,
              specular_threshold=0.0)
    up = {k: torch.randn(H, W, c, device=dev) * 1e-3 for k, c in dict(rgb=3, dpt=1, acc=1, norm=3, aux=2, dist=1).items()}
    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    ts = {"build": [], "fwd": [], "bwd": []}
    for it in range(6):
        e0 = ev(); RT.optix_build(v, f); e1 = ev()
        fwd = RT.trace_forward(training=True, backend="optix", rebuild=False, **kw); e2 = ev()
        RT.trace_backward(fwd, up, backend="optix", **kw); e3 = ev()
        torch.cuda.synchronize()
        if it >= 2:
            ts["build"].append(e0.elapsed_time(e1)); ts["fwd"].append(e1.elapsed_time(e2)); ts["bwd"].append(e2.elapsed_time(e3))
    rep["timing_ms"] = {k: sorted(x)[len(x) // 2] for k, x in ts.items()}
    print("timing", rep["timing_ms"], flush=True)
with open(os.path.join(ROOT, "gpurun_out", "optix_probe.json"), "w") as fjs:
    json.dump(rep, fjs, indent=1)
