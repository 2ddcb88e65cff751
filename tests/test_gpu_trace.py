"""GPU parity of the surfel tracer: LBVH CUDA path (through SurfelTracer, i.e. the C ABI) vs the brute-force float64 oracle.
Tolerances as in test_gpu_raster.py."""
import pytest
import torch

from envgs_b200 import scene
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4
GRAD_TOL = 1e-3


def _up(H, W, seed):
    return dict(rgb=Hh.upstream((H, W, 3), seed), dpt=Hh.upstream((H, W, 1), seed + 1, 0.1), acc=Hh.upstream((H, W, 1), seed + 2),
                norm=Hh.upstream((H, W, 3), seed + 3), aux=Hh.upstream((H, W, 2), seed + 4))


def _check(cu, orc, keys, frac_tol=1e-3, max_tol=3e-2, grad_tol=GRAD_TOL):
    for k in ("rgb", "acc", "norm", "aux", "dpt", "dist"):
        a, b = cu[k], orc[k]
        scale = max(1.0, float(b.abs().max()))
        assert Hh.frac_above(a, b, RGB_TOL * scale) <= frac_tol, Hh.image_report(k, a, b)
        assert float((a - b).abs().max()) <= max_tol * scale, Hh.image_report(k, a, b)
    assert Hh.rel_l2(cu["wet"].reshape(-1), orc["wet"].reshape(-1)) < 1e-3
    for k in keys:
        t, full = Hh.trimmed_rel_l2(cu[k], orc[k]), Hh.rel_l2(cu[k], orc[k])
        assert t <= grad_tol, f"{k}: trimmed rel L2 {t:.3e} (full {full:.3e})"
        assert full <= 10 * grad_tol, f"{k}: rel L2 {full:.3e}"


def test_reflection_trace_sh(built_lib):
    """EnvGS's call: start_from_first=False, max_trace_depth=0, SH colours evaluated along the ray."""
    env = scene.make_env_set(4000, seed=1, s0=0.04)
    cam = scene.ring_cameras(1, 48, 64, seed=2)[0]
    ro, rd = Hh.synthetic_rays(48, 64, seed=5)
    up = _up(48, 64, 20)
    cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, grads=up)
    _check(cu, orc, ["g_means", "g_opac", "g_scales", "g_rots", "g_shs", "g_ray_o", "g_ray_d", "g_grads3D"])
    # mid_val layout (config.h:18-26): ray, depth, alpha, normal, aux, rgb of segment 0
    assert Hh.frac_above(cu["mid"][..., :16], orc["mid"][..., :16], 1e-4 * max(1.0, float(orc["mid"].abs().max()))) < 1e-3


def test_dense_scene_many_chunks(built_lib):
    """Low opacity + many overlapping surfels: rays need several 16-hit chunks and reach the 256-contributor cap region."""
    env = scene.make_env_set(6000, seed=3, s0=0.2)
    env.opacities.mul_(0.05).add_(0.004)
    cam = scene.ring_cameras(1, 24, 32, seed=2)[0]
    ro, rd = Hh.synthetic_rays(24, 32, seed=6)
    up = _up(24, 32, 30)
    cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, grads=up, max_hits=1024, ray_chunk=256)
    _check(cu, orc, ["g_means", "g_opac", "g_scales", "g_rots", "g_shs", "g_ray_d"], frac_tol=3e-3, grad_tol=2e-3)


def test_camera_rays_start_from_first(built_lib):
    """Primary-ray mode (use_base_tracing): low-pass filter, near plane 0.2, acc gradient, precomputed colours + aux."""
    env = scene.make_base_set(3000, seed=5, s0=0.04)
    cam = scene.ring_cameras(1, 40, 56, seed=7)[0]
    ro, rd = scene.camera_rays(cam)
    up = _up(40, 56, 40)
    cu = Hh.run_cuda_trace(env, cam, ro, rd, use_sh=False, others=True, start_from_first=True, grads=up)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, use_sh=False, others=True, start_from_first=True, grads=up)
    _check(cu, orc, ["g_means", "g_opac", "g_scales", "g_rots", "g_colors", "g_others"])


def test_two_segments(built_lib):
    """max_trace_depth = 1: second bounce off the accumulated normal, specular chain in rgb and in the gradients."""
    env = scene.make_env_set(3000, seed=9, s0=0.08)
    env.specular.fill_(0.6)
    cam = scene.ring_cameras(1, 32, 40, seed=2)[0]
    ro, rd = Hh.synthetic_rays(32, 40, seed=8)
    up = _up(32, 40, 50)
    kw = dict(use_sh=True, others=True, max_trace_depth=1, specular_threshold=0.01, grads=up)
    cu = Hh.run_cuda_trace(env, cam, ro, rd, **kw)
    orc = Hh.run_oracle_trace(env, cam, ro, rd, **kw)
    _check(cu, orc, ["g_means", "g_opac", "g_shs", "g_others", "g_ray_d"], frac_tol=3e-3, grad_tol=3e-3)
    assert Hh.frac_above(cu["mid"], orc["mid"], 1e-4 * max(1.0, float(orc["mid"].abs().max()))) < 3e-3


def test_degenerate_rays_and_refit(built_lib):
    from envgs_b200.tracer import SurfelTracer, SurfelTracingSettings
    env = scene.make_env_set(2000, seed=2, s0=0.05).to("cuda")
    cam = scene.ring_cameras(1, 16, 16)[0].to("cuda")
    s = SurfelTracingSettings(16, 16, cam.tanfovx, cam.tanfovy, torch.zeros(3, device="cuda"), 1.0, cam.viewmatrix, cam.projmatrix,
                              3, cam.campos, False, False, 0, 0.0)
    tr = SurfelTracer().cuda()
    v, f = scene.quad_vertices(env.means, env.scales, env.rotations)
    tr.build_acceleration_structure(v, f, rebuild=True)
    gen = torch.Generator(device="cpu").manual_seed(123)
    ro = torch.zeros(16, 16, 3, device="cuda")
    rd = torch.randn(16, 16, 3, generator=gen).cuda()
    rd[0, 0] = 0.0                       # zero direction
    rd[0, 1] = float("nan")              # NaN direction
    rd[0, 2] = torch.tensor([1.0, 0.0, 0.0])   # axis-aligned
    out = tr(ro, rd, v, env.means, torch.zeros_like(env.means), shs=env.shs, opacities=env.opacities, scales=env.scales,
             rotations=env.rotations, tracer_settings=s, start_from_first=False)
    rgb = out[0]
    assert torch.isfinite(rgb[1:]).all() and float(out[2][0, 0]) == 0.0
    # refit (rebuild=0) after moving the surfels must equal a fresh build
    means2 = env.means + 0.05 * torch.randn(env.means.shape, generator=gen).cuda()
    v2, f2 = scene.quad_vertices(means2, env.scales, env.rotations)
    rd = torch.randn(16, 16, 3, generator=gen).cuda()
    tr.build_acceleration_structure(v2, f2, rebuild=False)
    a = tr(ro, rd, v2, means2, torch.zeros_like(means2), shs=env.shs, opacities=env.opacities, scales=env.scales,
           rotations=env.rotations, tracer_settings=s, start_from_first=False)
    tr.build_acceleration_structure(v2, f2, rebuild=True)
    b = tr(ro, rd, v2, means2, torch.zeros_like(means2), shs=env.shs, opacities=env.opacities, scales=env.scales,
           rotations=env.rotations, tracer_settings=s, start_from_first=False)
    assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[1], b[1], atol=1e-5)
    # tracing a surfel count the BVH was not built for is an error, not a silent wrong answer
    with pytest.raises(RuntimeError):
        tr(ro, rd, v2, means2[:100], torch.zeros(100, 3, device="cuda"), shs=env.shs[:100], opacities=env.opacities[:100],
           scales=env.scales[:100], rotations=env.rotations[:100], tracer_settings=s, start_from_first=False)


def test_full_size_properties(built_lib):
    """Config-3 sized trace (256K env surfels, 1080p rays): size-independent properties only."""
    env = scene.make_env_set(262_144, seed=1, s0=0.005)
    cam = scene.ring_cameras(1, 1080, 1920, seed=2)[0]
    ro, rd = Hh.synthetic_rays(1080, 1920, seed=5)
    up = dict(rgb=Hh.upstream((1080, 1920, 3), 1, 1e-3), dpt=torch.zeros(1080, 1920, 1), acc=torch.zeros(1080, 1920, 1),
              norm=torch.zeros(1080, 1920, 3), aux=torch.zeros(1080, 1920, 2))
    cu = Hh.run_cuda_trace(env, cam, ro, rd, grads=up)
    assert torch.isfinite(cu["rgb"]).all() and float(cu["acc"].min()) >= 0 and float(cu["acc"].max()) <= 1.0
    # per-surfel weights partition the per-ray alpha
    assert abs(float(cu["wet"].double().sum()) - float(cu["acc"].double().sum())) / float(cu["acc"].double().sum()) < 1e-4
    # colour gradient of SH band 0 is SH_C0 * sum_rays w * dL/drgb  ->  linear in the upstream gradient
    up2 = {k: 2 * v for k, v in up.items()}
    cu2 = Hh.run_cuda_trace(env, cam, ro, rd, grads=up2)
    assert Hh.rel_l2(cu2["g_shs"], 2 * cu["g_shs"]) < 1e-4 and Hh.rel_l2(cu2["g_means"], 2 * cu["g_means"]) < 1e-3
    # surfels nobody hit get no gradient
    untouched = cu["wet"][:, 0] == 0
    assert float(cu["g_means"][untouched].abs().max()) == 0


def test_fast_path_equals_general_kernel(built_lib, monkeypatch):
    """EnvGS' call shape takes the BVH4 + hit-parallel fast path; ENVGS_TRACE_FORCE_MONOLITHIC=1 routes the same call through
    the general chunk-loop kernel.  Both must agree (they share nothing but the BVH build)."""
    env = scene.make_env_set(5000, seed=4, s0=0.06)
    cam = scene.ring_cameras(1, 36, 52, seed=2)[0]
    ro, rd = Hh.synthetic_rays(36, 52, seed=9)
    up = _up(36, 52, 60)
    fast = Hh.run_cuda_trace(env, cam, ro, rd, grads=up)
    monkeypatch.setenv("ENVGS_TRACE_FORCE_MONOLITHIC", "1")
    gen = Hh.run_cuda_trace(env, cam, ro, rd, grads=up)
    monkeypatch.delenv("ENVGS_TRACE_FORCE_MONOLITHIC")
    for k in ("rgb", "dpt", "acc", "norm", "mid", "dist"):
        assert Hh.frac_above(fast[k], gen[k], 1e-5 * max(1.0, float(gen[k].abs().max()))) < 1e-4, Hh.image_report(k, fast[k], gen[k], 1e-5)
    for k in ("g_means", "g_opac", "g_scales", "g_rots", "g_shs", "g_ray_o", "g_ray_d"):
        assert Hh.rel_l2(fast[k], gen[k]) < 2e-4, f"{k}: {Hh.rel_l2(fast[k], gen[k]):.3e}"
    assert Hh.rel_l2(fast["wet"].reshape(-1), gen["wet"].reshape(-1)) < 1e-5


def test_two_segment_fast_path_equals_general_kernel(built_lib, monkeypatch):
    """max_trace_depth = 1 runs segment by segment on the hit-parallel kernels (trace_fast.cu, trace_forward_fast); the same call
    through the general kernel (ENVGS_TRACE_FORCE_MONOLITHIC=1), which walks both segments of a ray in one thread, must agree
    — including rays that stop after the first segment and the dense rays that overflow into the 16-hit chunk loop."""
    env = scene.make_env_set(6000, seed=9, s0=0.09)
    env.specular.copy_(torch.rand(env.specular.shape, generator=torch.Generator().manual_seed(1)) * 0.9)   # some below the threshold
    cam = scene.ring_cameras(1, 36, 52, seed=2)[0]
    ro, rd = Hh.synthetic_rays(36, 52, seed=8)
    up = _up(36, 52, 70)
    kw = dict(use_sh=True, others=True, max_trace_depth=1, specular_threshold=0.3, grads=up)
    fast = Hh.run_cuda_trace(env, cam, ro, rd, **kw)
    monkeypatch.setenv("ENVGS_TRACE_FORCE_MONOLITHIC", "1")
    gen = Hh.run_cuda_trace(env, cam, ro, rd, **kw)
    monkeypatch.delenv("ENVGS_TRACE_FORCE_MONOLITHIC")
    stopped = gen["mid"][..., 16:].abs().sum(-1) == 0
    assert 0.02 < float(stopped.float().mean()) < 0.98          # both kinds of ray are present
    assert torch.equal(stopped, fast["mid"][..., 16:].abs().sum(-1) == 0)
    for k in ("rgb", "dpt", "acc", "norm", "aux", "mid", "dist"):
        assert Hh.frac_above(fast[k], gen[k], 1e-5 * max(1.0, float(gen[k].abs().max()))) < 1e-4, Hh.image_report(k, fast[k], gen[k], 1e-5)
    for k in ("g_means", "g_opac", "g_scales", "g_rots", "g_shs", "g_others", "g_ray_o", "g_ray_d"):
        assert Hh.rel_l2(fast[k], gen[k]) < 2e-4, f"{k}: {Hh.rel_l2(fast[k], gen[k]):.3e}"
    assert Hh.rel_l2(fast["wet"].reshape(-1), gen["wet"].reshape(-1)) < 1e-5
