"""Golden vectors for the host-side restatements around the hot path, from the UNMODIFIED reference functions (CPU, build container).

Test infrastructure only.  Each entry of tests/golden/host_ref.pt is the output of a reference function on seeded inputs:
  * `prepare_gaussian_camera` (easyvolcap/utils/gaussian2d_utils.py:67-100)  <- envgs_b200.gaussian_io.camera_from_batch
  * `dpt2norm` / `dpt2xyz`   (gaussian2d_utils.py:1158-1206)                 <- envgs_b200.pipeline.depth_to_normal
  * `eval_sh` (degree 3)      (easyvolcap/utils/sh_utils.py:643-700), called as `render()` does (gaussian2d_utils.py:1071-1084:
    coefficients (P,3,16), normalised `xyz - camera_center`)                  <- envgs_b200.pipeline.eval_sh_deg3 (the PyTorch twin
    of the fused feature kernel, tests/test_gpu_raster.py::test_fused_sh_features_match_pytorch)
  * `HardwareRendering.get_disks` (easyvolcap/utils/optix_utils.py:39-69: surfel -> 4 quad corners + 2 triangles, the tracer's
    vertex buffer)                                                             <- envgs_b200.scene.quad_vertices (twin of glue.quad_vertices)
Only the reference files themselves are executed (helper imports stubbed, see oracle/make_densify_golden.py).

    python oracle/make_host_golden.py        # writes tests/golden/host_ref.pt
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_densify_golden import load_reference_module  # noqa: E402

SH_FILE = "/root/reference/easyvolcap/utils/sh_utils.py"
OPTIX_FILE = "/root/reference/easyvolcap/utils/optix_utils.py"
H, W = 54, 96


def inputs():
    from envgs_b200 import scene
    g = torch.Generator().manual_seed(17)
    cam = scene.ring_cameras(3, H, W, seed=6)[2]
    depth = (2.5 + torch.rand(1, H, W, generator=g) * 2.0).float()
    P = 500
    shs = (torch.randn(P, 16, 3, generator=g) * 0.3).float()
    xyz = (torch.randn(P, 3, generator=g)).float()
    return cam, depth, shs, xyz


def disk_inputs():
    from envgs_b200 import scene
    return scene.make_env_set(300, seed=8, s0=0.05)


def run():
    G, dotdict = load_reference_module()
    S, _ = load_reference_module(SH_FILE, "ref_sh_utils")
    cam, depth, shs, xyz = inputs()
    n, f = torch.tensor(cam.znear), torch.tensor(cam.zfar)
    one = lambda t: t[None]
    meta = dotdict(H=torch.tensor([H]), W=torch.tensor([W]), K=one(cam.K), R=one(cam.R), T=one(cam.T), n=one(n), f=one(f))
    batch = dotdict(H=torch.tensor([H]), W=torch.tensor([W]), K=one(cam.K), R=one(cam.R), T=one(cam.T), n=one(n), f=one(f), meta=meta)
    c = G.prepare_gaussian_camera(batch)
    out = {"camera": {"world_view_transform": c.world_view_transform.clone(), "full_proj_transform": c.full_proj_transform.clone(),
                      "camera_center": c.camera_center.clone(), "tanfovx": float(c.tanfovx), "tanfovy": float(c.tanfovy)}}
    out["dpt2norm"] = G.dpt2norm(c, depth, device="cpu").clone()
    dirs = torch.nn.functional.normalize(xyz - c.camera_center[None], dim=-1)
    out["eval_sh_deg3"] = S.eval_sh(3, shs.transpose(1, 2), dirs).clone()          # (P,3,16) layout, as render() builds it
    # quad corners: a reference GaussianModel holding the seeded set (pre-activation parameters), then the unbound get_disks
    import types
    env = disk_inputs()
    model = G.GaussianModel(xyz=env.means.clone(), colors=None, init_scale=torch.log(env.scales), xyz_lr_scheduler=None)
    model._xyz.data, model._scaling.data, model._rotation.data = env.means.clone(), torch.log(env.scales), env.rotations.clone()
    mk = lambda name, **attrs: (lambda m: (m.__dict__.update(attrs), m)[1])(types.ModuleType(name))
    stubs = {"diff_surfel_tracing": mk("diff_surfel_tracing", SurfelTracer=object, SurfelTracingSettings=object),
             "easyvolcap.utils.base_utils": mk("easyvolcap.utils.base_utils", dotdict=dotdict),
             "easyvolcap.utils.math_utils": mk("easyvolcap.utils.math_utils", normalize=torch.nn.functional.normalize),
             "easyvolcap.utils.nerf_utils": mk("easyvolcap.utils.nerf_utils", volume_rendering=None),
             "easyvolcap.utils.gaussian2d_utils": G}
    O, _ = load_reference_module(OPTIX_FILE, "ref_optix_utils", extra_stubs=stubs)
    with torch.no_grad():
        v, f = O.HardwareRendering.get_disks(None, model)
    out["get_disks"] = {"v": v.clone(), "f": f.clone()}
    out["reference"] = "zju3dv/EnvGS easyvolcap/utils/{gaussian2d_utils,sh_utils,optix_utils}.py, unmodified, CPU, torch " + torch.__version__
    return out


if __name__ == "__main__":
    res = run()
    torch.save(res, os.path.join(ROOT, "tests", "golden", "host_ref.pt"))
    print({k: tuple(v.shape) for k, v in res.items() if torch.is_tensor(v)}, tuple(res["get_disks"]["v"].shape), res["camera"]["tanfovx"])
