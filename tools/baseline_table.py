"""Fill BASELINE.md §2c for the single-GPU rasterizer configs (GPU box): UNMODIFIED reference rasterizer (oracle/_ref, sm_100a build)
vs ours through the same operator call, in-kernel SH degree 3 (C = 3).  Config 1 also times the CPU oracle (forward)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from envgs_b200.rasterizer import GaussianRasterizationSettings, make_module
from oracle import build_ref
from tests import helpers as Hh

def timed(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def run(P, H, W, bwd, cpu):
    dev = torch.device("cuda")
    s0 = 0.005 if H >= 1080 else 0.005 * 1080 / H
    base = scene.make_base_set(P, seed=0, s0=s0)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    gc, go = Hh.upstream((3, H, W), 3, 1e-3).to(dev), Hh.upstream((7, H, W), 4, 1e-3).to(dev)
    _C = build_ref.load("ref_rast_wet")
    e = torch.empty(0, device=dev)
    means, opac, scales, rots, shs = [t.to(dev) for t in (base.means, base.opacities, base.scales, base.rotations, base.shs)]
    vm, pm, cp, bg = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev), torch.zeros(3, device=dev)
    st = {}
    def ref_f(): st["f"] = _C.rasterize_gaussians(bg, means, e, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, shs, 3, cp, False, False)
    def ref_b():
        n, color, allmap, weight, radii, gb, bb, ib = st["f"]
        st["b"] = _C.rasterize_gaussians_backward(bg, means, radii, e, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, gc, go, shs, 3, cp, gb, n, bb, ib, False)
    _, _, Rast = make_module(3)
    s = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, vm, pm, 3, cp, False, False)
    leafs = [t.clone().requires_grad_(True) for t in (means, opac, scales, rots, shs)]
    m2 = torch.zeros_like(means, requires_grad=True)
    def our_f(): st["o"] = Rast(s)(means3D=leafs[0], means2D=m2, opacities=leafs[1], shs=leafs[4], scales=leafs[2], rotations=leafs[3])
    def our_b():
        color, radii, allmap, weight = st["o"]
        for l in leafs + [m2]: l.grad = None
        torch.autograd.backward([color, allmap], [gc, go], retain_graph=True)
    r = {"P": P, "H": H, "W": W, "ref_fwd_ms": timed(ref_f), "ours_fwd_ms": timed(our_f)}
    if bwd:
        r["ref_bwd_ms"], r["ours_bwd_ms"] = timed(ref_b), timed(our_b)
    tot_r = r["ref_fwd_ms"] + r.get("ref_bwd_ms", 0.0); tot_o = r["ours_fwd_ms"] + r.get("ours_bwd_ms", 0.0)
    r["ref_Mpix_s"], r["ours_Mpix_s"], r["speedup"] = H * W / tot_r / 1e3, H * W / tot_o / 1e3, tot_r / tot_o
    r["color_max_abs_diff"] = float((st["o"][0] - st["f"][1]).abs().max())
    if cpu:
        from oracle import raster_oracle as RO
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        t = time.perf_counter()
        with torch.no_grad():
          RO.rasterize(base.means.float(), base.opacities.float(), base.scales.float(), base.rotations.float(), shs=base.shs.float(), sh_degree=3,
                     viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos, bg=torch.zeros(3), W=W, H=H)
        r["cpu_oracle_fwd_s"], r["cpu_threads"] = time.perf_counter() - t, torch.get_num_threads()
        r["cpu_Mpix_s"] = H * W / r["cpu_oracle_fwd_s"] / 1e6
    return r

if __name__ == "__main__":
    out = {"config1 (10K base, 256x256, SH3, forward)": run(10_000, 256, 256, False, True),
           "config2 (500K base, 1080p, SH3, forward+backward)": run(500_000, 1080, 1920, True, False)}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "baseline_table.json"), "w"), indent=1)
