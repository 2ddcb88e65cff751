"""Development aid (GPU box): bit-level comparison of the per-surfel preprocess state (homography T, projected centre,
radius) between this library and the UNMODIFIED reference build (oracle/_ref), plus the reference's own sensitivity to a
1-ulp perturbation of its inputs (the conditioning floor of the gradient comparison)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from oracle import build_ref
from tests import helpers as Hh


def ref_geom_views(geom: torch.Tensor, P: int):
    """Carve the reference's GeometryState (rasterizer_impl.cu:155-170): depths, clamped, radii, means2D, transMat."""
    base = geom.data_ptr()
    def take(off, nbytes):
        off = (base + off + 127) // 128 * 128 - base
        return off, off + nbytes
    o, e = take(0, 4 * P); depths = geom[o:e].view(torch.float32)
    o, e = take(e, 3 * P)
    o, e = take(e, 4 * P); radii = geom[o:e].view(torch.int32)
    o, e = take(e, 8 * P); means2D = geom[o:e].view(torch.float32).view(P, 2)
    o, e = take(e, 36 * P); T = geom[o:e].view(torch.float32).view(P, 9)
    return depths, radii, means2D, T


def main(P=1_000_000, H=1080, W=1920, C=5):
    dev = torch.device("cuda")
    base = scene.make_base_set(P, seed=0, s0=0.005)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    feats = Hh.base_features(base, cam, C)
    gc, go = Hh.upstream((C, H, W), 3, 1e-3), Hh.upstream((7, H, W), 4, 1e-3)
    _C = build_ref.load("ref_rast_wet_ch05")
    e = torch.empty(0, device=dev)
    means, opac, scales, rots, colors = [t.to(dev) for t in (base.means, base.opacities, base.scales, base.rotations, feats)]
    vm, pm, cp = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
    bgd = torch.zeros(C, device=dev)

    def ref(means, scales, rots):
        n, color, allmap, weight, radii, gb, bb, ib = _C.rasterize_gaussians(bgd, means, colors, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, e, 3, cp, False, False)
        g = _C.rasterize_gaussians_backward(bgd, means, radii, colors, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, gc.to(dev), go.to(dev), e, 3, cp, gb, n, bb, ib, False)
        return dict(n=n, color=color, allmap=allmap, radii=radii, gb=gb, g_means2D=g[0], g_colors=g[1], g_opac=g[2], g_means=g[3], g_scales=g[6], g_rots=g[7])

    r0 = ref(means, scales, rots)
    _, _, r_m2d, r_T = ref_geom_views(r0["gb"], P)
    r_radii = r0["radii"]   # the op is handed a radii tensor, so GeometryState::internal_radii stays unused

    from envgs_b200.rasterizer import GaussianRasterizationSettings, make_module
    _, _, Rast = make_module(C)
    s = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bgd, 1.0, vm, pm, 3, cp, False, False)
    leafs = [t.clone().requires_grad_(True) for t in (means, opac, scales, rots, colors)]
    m2 = torch.zeros_like(means, requires_grad=True)
    color, radii, allmap, weight = Rast(s)(means3D=leafs[0], means2D=m2, opacities=leafs[1], colors_precomp=leafs[4], scales=leafs[2], rotations=leafs[3])
    geom = color.grad_fn.saved_tensors[8]
    off = (geom.data_ptr() + 255) // 256 * 256 - geom.data_ptr()
    rec = geom[off:off + 64 * P].view(torch.float32).view(P, 16)
    vis = r_radii > 0
    T_ours = rec[:, :9]
    rep = {"P": P, "visible_ref": int(vis.sum())}
    rep["T_bit_mismatch_surfels"] = int((T_ours[vis].view(torch.int32) != r_T[vis].view(torch.int32)).any(-1).sum())
    rep["T_bit_mismatch_by_entry"] = (T_ours[vis].view(torch.int32) != r_T[vis].view(torch.int32)).sum(0).tolist()
    rep["center_bit_mismatch"] = int((rec[vis][:, 9:11].contiguous().view(torch.int32) != r_m2d[vis].contiguous().view(torch.int32)).any(-1).sum())
    rep["radii_mismatch"] = int((radii != r_radii).sum())
    dc = (color.detach() - r0["color"]).abs()
    rep["color_max"] = float(dc.max()); rep["color_frac_gt_1e-4"] = float((dc > 1e-4).float().mean()); rep["color_pixels_gt_1e-4"] = int((dc > 1e-4).any(0).sum())
    sc = r0["allmap"].abs().amax(dim=(1, 2), keepdim=True).clamp_min(1.0)
    da = ((allmap.detach() - r0["allmap"]).abs() / sc)
    rep["allmap_max_scaled"] = float(da.max()); rep["allmap_frac_gt_1e-4"] = float((da > 1e-4).float().mean())
    torch.autograd.backward([color, allmap], [gc.to(dev), go.to(dev)])
    ours = dict(g_means=leafs[0].grad, g_opac=leafs[1].grad, g_scales=leafs[2].grad, g_rots=leafs[3].grad, g_colors=leafs[4].grad, g_means2D=m2.grad)
    for k, v in ours.items():
        rep[f"{k}_relL2"] = Hh.rel_l2(v.cpu(), r0[k].cpu())
    # conditioning floor: the reference against itself with every input nudged by one ulp (seeded sign)
    g = torch.Generator(device="cpu").manual_seed(99)
    def nudge(t):
        sgn = (torch.randint(0, 2, t.shape, generator=g) * 2 - 1).to(dev)
        return torch.where(sgn > 0, torch.nextafter(t, t + 1), torch.nextafter(t, t - 1))
    r1 = ref(nudge(means), nudge(scales), nudge(rots))
    for k in ours:
        rep[f"{k}_ref_vs_ref_1ulp"] = Hh.rel_l2(r1[k].cpu(), r0[k].cpu())
    rep["color_ref_vs_ref_1ulp_max"] = float((r1["color"] - r0["color"]).abs().max())
    rep["radii_ref_vs_ref_1ulp_mismatch"] = int((r1["radii"] != r0["radii"]).sum())
    # who carries the error
    d = (ours["g_means"] - r0["g_means"]).norm(dim=1)
    top = d.topk(10)
    Tw = r_T[:, 6:9]
    dd = 9 * Tw[:, 0] ** 2 + 9 * Tw[:, 1] ** 2 - Tw[:, 2] ** 2
    rep["worst_surfels"] = [dict(id=int(i), err=float(v), ref_norm=float(r0["g_means"][i].norm()), d_over_Twz2=float(dd[i] / Tw[i, 2] ** 2), radius=int(r_radii[i])) for v, i in zip(top.values, top.indices)]
    rep["g_means_total_norm"] = float(r0["g_means"].norm())
    print(json.dumps(rep, indent=1))
    tag = os.environ.get("REF_BITS_TAG", "default")
    with open(os.path.join(ROOT, "gpurun_out", f"ref_bits_P{P}_{tag}.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main(P=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
