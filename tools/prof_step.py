"""Profiling driver (GPU box): N forward+backward steps of the config-3 pipeline without the optimizer.
Used under ncu (see profiles/README.md) and for the traversal statistics."""
import os, sys, json, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene, _lib
from envgs_b200.pipeline import EnvGSRenderer
import bench

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "config3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    stats = "--stats" in sys.argv
    Pb, Pe, H, W = bench.WORKLOADS[wl]
    dev = torch.device("cuda")
    s0 = 0.005 * 1080 / H
    base = scene.make_base_set(Pb, seed=0, s0=s0)
    env = scene.make_env_set(Pe, seed=1, s0=s0) if Pe else None
    r = EnvGSRenderer(base, env, device=dev)
    cams = [c.to(dev) for c in scene.ring_cameras(4, H, W, seed=2)]
    target = torch.rand(H, W, 3, device=dev)
    if stats and Pe:
        with torch.no_grad():
            r(cams[0], target)            # builds the BVH (the native handle is created lazily on the first build)
        _lib.lib().envgs_debug_trace_stats(r.tracer.optix_context.handle, 1, None)
    for s in range(steps):
        for p in r.parameters(): p.grad = None
        out = r(cams[s % 4], target)
        r.loss(out, target).backward()
    torch.cuda.synchronize()
    if stats and Pe:
        buf = (ctypes.c_uint64 * 8)()
        _lib.lib().envgs_debug_trace_stats(r.tracer.optix_context.handle, 0, buf)
        rays, nodes, prims, hits, chunks, contrib = [int(buf[i]) for i in range(6)]
        print(json.dumps({"ray_segments(fwd+bwd)": rays, "nodes/ray": nodes / rays, "quad_tests/ray": prims / rays, "hits_inserted/ray": hits / rays,
                          "chunks/ray": chunks / rays, "contributions/ray": contrib / rays}))
        acc = out["env"]["acc"]
        print("env acc mean", float(acc.mean()), "alpha mean", float(out["allmap"][1].mean()))

if __name__ == "__main__":
    main()
