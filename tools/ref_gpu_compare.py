"""Development/measurement aid (GPU box): time the UNMODIFIED reference rasterizer (oracle/_ref, sm_100a build) against ours
on the benchmark scene and compare outputs.  Not part of the product path."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene, _lib
from oracle import build_ref
from tests import helpers as Hh

def timed(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

def main(P=1_000_000, H=1080, W=1920, C=5):
    dev = torch.device("cuda")
    base = scene.make_base_set(P, seed=0, s0=0.005)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    feats = Hh.base_features(base, cam, C)
    gc, go = Hh.upstream((C, H, W), 3, 1e-3), Hh.upstream((7, H, W), 4, 1e-3)
    bg = torch.zeros(C)
    _C = build_ref.load("ref_rast_wet_ch05")
    e = torch.empty(0, device=dev)
    means, opac, scales, rots, colors = [t.to(dev) for t in (base.means, base.opacities, base.scales, base.rotations, feats)]
    vm, pm, cp = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
    gcd, god, bgd = gc.to(dev), go.to(dev), bg.to(dev)
    state = {}
    def ref_fwd():
        state["f"] = _C.rasterize_gaussians(bgd, means, colors, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, e, 3, cp, False, False)
    def ref_bwd():
        n, color, allmap, weight, radii, gb, bb, ib = state["f"]
        state["b"] = _C.rasterize_gaussians_backward(bgd, means, radii, colors, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, gcd, god, e, 3, cp, gb, n, bb, ib, False)
    t_rf = timed(ref_fwd); t_rb = timed(ref_bwd)
    n_ref = state["f"][0]
    # ours through the operator
    from envgs_b200.rasterizer import GaussianRasterizationSettings, make_module
    _, _, Rast = make_module(C)
    s = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bgd, 1.0, vm, pm, 3, cp, False, False)
    leafs = [t.clone().requires_grad_(True) for t in (means, opac, scales, rots, colors)]
    m2 = torch.zeros_like(means, requires_grad=True)
    ost = {}
    def our_fwd():
        ost["o"] = Rast(s)(means3D=leafs[0], means2D=m2, opacities=leafs[1], colors_precomp=leafs[4], scales=leafs[2], rotations=leafs[3])
    def our_bwd():
        color, radii, allmap, weight = ost["o"]
        for l in leafs + [m2]: l.grad = None
        torch.autograd.backward([color, allmap], [gcd, god], retain_graph=True)
    t_of = timed(our_fwd); t_ob = timed(our_bwd)
    _lib.profile_enable(True); our_fwd(); our_bwd(); torch.cuda.synchronize(); prof = _lib.profile_report(); _lib.profile_enable(False)
    color, radii, allmap, weight = ost["o"]
    rn, rcolor, rallmap, rweight, rradii = state["f"][:5]
    gm2, gcol, gop, gm3, gcov, gsh, gsc, grot = state["b"]
    rep = {
        "P": P, "HxW": [H, W], "ref_fwd_ms": t_rf, "ref_bwd_ms": t_rb, "ours_fwd_ms": t_of, "ours_bwd_ms": t_ob,
        "speedup_fwd": t_rf / t_of, "speedup_bwd": t_rb / t_ob, "speedup_total": (t_rf + t_rb) / (t_of + t_ob),
        "ref_num_rendered": int(rn), "ours_kernels_ms": {k: v[1] / v[0] for k, v in prof.items()},
        "color": Hh.image_report("color", color.cpu(), rcolor.cpu()), "allmap": Hh.image_report("allmap", allmap.detach().cpu(), rallmap.cpu()),
        "radii_mismatch": int((radii != rradii).sum()), "radii_max_abs_diff": int((radii.long() - rradii.long()).abs().max()),
        "g_means_trimmed": Hh.trimmed_rel_l2(leafs[0].grad.cpu(), gm3.cpu()), "g_rots_trimmed": Hh.trimmed_rel_l2(leafs[3].grad.cpu(), grot.cpu()),
        "g_scales_trimmed": Hh.trimmed_rel_l2(leafs[2].grad.cpu(), gsc.cpu()), "g_means2D_trimmed": Hh.trimmed_rel_l2(m2.grad.cpu(), gm2.cpu()),
        "weight_relL2": Hh.rel_l2(weight.cpu(), rweight.cpu()),
        "g_means": Hh.rel_l2(leafs[0].grad.cpu(), gm3.cpu()), "g_opac": Hh.rel_l2(leafs[1].grad.cpu(), gop.cpu()),
        "g_scales": Hh.rel_l2(leafs[2].grad.cpu(), gsc.cpu()), "g_rots": Hh.rel_l2(leafs[3].grad.cpu(), grot.cpu()),
        "g_colors": Hh.rel_l2(leafs[4].grad.cpu(), gcol.cpu()), "g_means2D": Hh.rel_l2(m2.grad.cpu(), gm2.cpu()),
    }
    # reference run-to-run noise
    ref_bwd(); g2 = state["b"][3]
    rep["ref_noise_g_means"] = Hh.rel_l2(g2.cpu(), gm3.cpu())
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ref_gpu_compare_P{P}.json"), "w") as f:
        json.dump(rep, f, indent=1)

if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    main(P=P)
