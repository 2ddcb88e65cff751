"""Print CUDA-vs-oracle error metrics for a set of small cases (development aid; run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from envgs_b200 import scene
from tests import helpers as Hh

def raster_case(P, H, W, C, s0, seed, use_sh=False):
    base = scene.make_base_set(P, seed=seed, s0=s0)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    feats = None if use_sh else Hh.base_features(base, cam, C)
    gc, go = Hh.upstream((C, H, W), 3), Hh.upstream((7, H, W), 4)
    cu = Hh.run_cuda_raster(base, cam, C, feats=feats, use_sh=use_sh, g_color=gc, g_others=go)
    orc = Hh.run_oracle_raster(base, cam, C, feats=feats, use_sh=use_sh, g_color=gc, g_others=go)
    print(f"--- raster P={P} {H}x{W} C={C} sh={use_sh}")
    for k in ("color", "allmap"):
        print("  ", Hh.image_report(k, cu[k], orc[k]))
    print("   radii mismatch", int((cu['radii'] != orc['radii']).sum()), "weight relL2", Hh.rel_l2(cu['weight'], orc['weight']))
    for k in cu:
        if k.startswith("g_"):
            d = (cu[k].double() - orc[k].double())
            # per-surfel error concentration
            per = d.reshape(d.shape[0], -1).norm(dim=1)
            top = per.topk(3).values / max(1e-30, float(orc[k].double().norm()))
            print(f"   {k}: relL2 {Hh.rel_l2(cu[k], orc[k]):.3e}  top3 per-surfel share {top.tolist()}")

def trace_case(P, H, W, s0, seed, **kw):
    env = scene.make_env_set(P, seed=seed, s0=s0)
    cam = scene.ring_cameras(1, H, W, seed=2)[0]
    ro, rd = Hh.synthetic_rays(H, W, seed=5)
    up = dict(rgb=Hh.upstream((H, W, 3), 20), dpt=Hh.upstream((H, W, 1), 21, 0.1), acc=Hh.upstream((H, W, 1), 22),
              norm=Hh.upstream((H, W, 3), 23), aux=Hh.upstream((H, W, 2), 24))
    cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up, **kw)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, grads=up, **kw)
    print(f"--- trace P={P} {H}x{W} {kw}")
    for k in ("rgb", "dpt", "acc", "norm", "aux", "dist", "mid"):
        print("  ", Hh.image_report(k, cu[k], orc[k]))
    print("   wet relL2", Hh.rel_l2(cu['wet'].reshape(-1), orc['wet'].reshape(-1)))
    for k in cu:
        if k.startswith("g_"):
            print(f"   {k}: relL2 {Hh.rel_l2(cu[k], orc[k]):.3e}")

if __name__ == "__main__":
    raster_case(3000, 96, 128, 3, 0.03, 0)
    raster_case(3000, 96, 128, 5, 0.03, 0)
    raster_case(2500, 80, 112, 3, 0.03, 7, use_sh=True)
    trace_case(4000, 48, 64, 0.04, 1)
    trace_case(3000, 32, 40, 0.08, 9, others=True, max_trace_depth=1, specular_threshold=0.01)
