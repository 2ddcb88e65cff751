import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from envgs_b200 import scene
from tests import helpers as Hh
P,H,W,C,s0,seed = 3000,96,128,5,0.03,0
base = scene.make_base_set(P, seed=seed, s0=s0)
cam = scene.ring_cameras(1, H, W, seed=2)[0]
feats = Hh.base_features(base, cam, C)
for mode in ("all", "nomedian", "only_color", "only_depth", "only_alpha", "only_normal", "only_median", "only_dist"):
    gc, go = Hh.upstream((C, H, W), 3), Hh.upstream((7, H, W), 4)
    if mode == "nomedian": go[5] = 0
    if mode.startswith("only"):
        keep = dict(only_color=None, only_depth=0, only_alpha=1, only_normal=(2,5), only_median=5, only_dist=6)[mode]
        g2 = torch.zeros_like(go)
        if keep is None: pass
        elif isinstance(keep, tuple): g2[keep[0]:keep[1]] = go[keep[0]:keep[1]]
        else: g2[keep] = go[keep]
        go = g2
        if mode != "only_color": gc = torch.zeros_like(gc)
    cu = Hh.run_cuda_raster(base, cam, C, feats=feats, g_color=gc, g_others=go)
    orc = Hh.run_oracle_raster(base, cam, C, feats=feats, g_color=gc, g_others=go)
    line = [mode]
    for k in ("g_means", "g_opac", "g_scales", "g_rots", "g_colors", "g_means2D"):
        line.append(f"{k}={Hh.rel_l2(cu[k], orc[k]):.2e}")
    print(" ".join(line))
    if mode == "all":
        d = (cu["g_means"].double() - orc["g_means"].double()).norm(dim=1)
        i = int(d.argmax())
        print("worst surfel", i, "cu", cu["g_means"][i].tolist(), "or", orc["g_means"][i].tolist(), "opac", float(base.opacities[i]), "scale", base.scales[i].tolist(), "radii", int(cu["radii"][i]))
        print("  g_scales", cu["g_scales"][i].tolist(), orc["g_scales"][i].tolist(), "g_m2d", cu["g_means2D"][i].tolist(), orc["g_means2D"][i].tolist())
