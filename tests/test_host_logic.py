"""CPU: host-side logic — deterministic scene generation, camera conventions, oracle self-consistency, and the
view-parallel gradient exchange on gloo with world_size 2."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from envgs_b200 import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scene_is_deterministic_and_well_formed():
    a, b = scene.make_base_set(5000, seed=0), scene.make_base_set(5000, seed=0)
    for f in a.__dataclass_fields__:
        assert torch.equal(getattr(a, f), getattr(b, f))
    assert a.means.shape == (5000, 3) and a.shs.shape == (5000, 16, 3)
    assert torch.allclose(a.rotations.norm(dim=-1), torch.ones(5000), atol=1e-5)
    assert float(a.opacities.min()) > 0 and float(a.opacities.max()) < 1 and float(a.scales.min()) > 0
    env = scene.make_env_set(1000, seed=1)
    r = env.means.norm(dim=-1)
    assert float(r.min()) >= 8 - 1e-4 and float(r.max()) <= 20 + 1e-4


def test_camera_conventions_match_easyvolcap():
    cam = scene.ring_cameras(3, 90, 160, seed=2)[1]
    # viewmatrix = W2C^T, projmatrix = (Proj W2C)^T, campos = -R^T T  (gaussian2d_utils.py:86-89)
    w2c = cam.viewmatrix.t()
    assert torch.allclose(w2c[:3, :3], cam.R, atol=1e-6) and torch.allclose(w2c[:3, 3:], cam.T, atol=1e-6)
    assert torch.allclose(cam.campos, (-cam.R.t() @ cam.T)[:, 0], atol=1e-5)
    p = torch.tensor([0.1, -0.2, 0.3, 1.0])
    clip = cam.projmatrix.t() @ p
    view = w2c @ p
    assert abs(float(clip[3] - view[2])) < 1e-5                       # clip w == view z
    ro, rd = scene.camera_rays(cam)
    d_cam = rd @ cam.R.t()
    assert torch.allclose(d_cam[..., 2], torch.ones(90, 160), atol=1e-4)   # z_depth rays: camera-space z = 1
    assert float((ro - cam.campos).abs().max()) < 1e-5


def test_quad_vertices_order_matches_get_disks():
    s = scene.make_env_set(10, seed=3)
    v, f = scene.quad_vertices(s.means, s.scales, s.rotations)
    assert v.shape == (40, 3) and f.shape == (20, 3) and f.dtype == torch.int32
    c = v.reshape(10, 4, 3).mean(1)
    assert torch.allclose(c, s.means, atol=1e-5)
    eu = (v.reshape(10, 4, 3)[:, 3] - v.reshape(10, 4, 3)[:, 1]).norm(dim=-1)      # (+,-) - (-,-) = 6 s_u
    assert torch.allclose(eu, 6 * s.scales[:, 0], rtol=1e-4)
    assert f[0].tolist() == [0, 1, 2] and f[1].tolist() == [1, 2, 3]


def test_oracle_linearity_and_partition_of_unity():
    from tests import helpers as Hh
    base = scene.make_base_set(400, seed=5, s0=0.05)
    cam = scene.ring_cameras(1, 32, 48, seed=1)[0]
    feats = Hh.base_features(base, cam, 5)
    gc, go = Hh.upstream((5, 32, 48), 1), Hh.upstream((7, 32, 48), 2)
    a = Hh.run_oracle_raster(base, cam, 5, feats=feats, g_color=gc, g_others=go)
    b = Hh.run_oracle_raster(base, cam, 5, feats=feats, g_color=2 * gc, g_others=2 * go)
    assert Hh.rel_l2(b["g_means"], 2 * a["g_means"]) < 1e-6            # backward is linear in the upstream gradient
    alpha = a["allmap"][1]
    assert abs(float(a["weight"].sum()) - float(alpha.sum())) < 1e-3 * float(alpha.sum())   # weights partition alpha
    assert float(alpha.max()) <= 1.0 and float(alpha.min()) >= 0.0


def test_trace_oracle_single_surfel_closed_form():
    """One surfel facing the ray: alpha = o * exp(-(u^2+v^2)/2), rgb = alpha * colour, depth = alpha * t."""
    from oracle import trace_oracle as TO
    means = torch.tensor([[0.0, 0.0, 5.0]], dtype=torch.float64)
    scales = torch.tensor([[0.5, 0.25]], dtype=torch.float64)
    rots = torch.tensor([[1.0, 0.0, 0.0, 0.0]], dtype=torch.float64)
    opac = torch.tensor([[0.8]], dtype=torch.float64)
    col = torch.tensor([[0.2, 0.4, 0.6]], dtype=torch.float64)
    ro = torch.zeros(1, 2, 3, dtype=torch.float64)
    rd = torch.tensor([[[0.05, 0.02, 1.0], [1.0, 0.0, 0.0]]], dtype=torch.float64)    # second ray misses
    out = TO.trace(ro, rd, means, opac, scales, rots, colors_precomp=col, campos=torch.zeros(3, dtype=torch.float64),
                   bg=torch.zeros(3, dtype=torch.float64))
    u, v = 0.25 / 0.5, 0.1 / 0.25
    alpha = 0.8 * torch.exp(torch.tensor(-(u * u + v * v) / 2, dtype=torch.float64))
    assert torch.allclose(out.acc[0, 0, 0], alpha, atol=1e-12)
    assert torch.allclose(out.rgb[0, 0], alpha * col[0], atol=1e-12)
    assert torch.allclose(out.dpt[0, 0, 0], alpha * 5.0, atol=1e-12)
    assert float(out.acc[0, 1, 0]) == 0.0 and float(out.wet[0, 0]) == pytest.approx(float(alpha))
    # the normal faces the ray origin side (flip by mu - campos, forward.cu:315-321)
    assert float(out.norm[0, 0, 2]) < 0


def test_shard_views_round_robin():
    from envgs_b200.dist import shard_views
    assert shard_views(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard_views(138, r, 8) for r in range(8)), [])) == list(range(138))


class _ToyRenderer(torch.nn.Module):
    """CPU stand-in with the renderer's interface (forward(cam) -> dict, loss(out, target)); `env_c` plays the environment
    set: created last, so its gradient is produced FIRST in the backward pass and its exchange is launched from the hook."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(7, 3, generator=g))
        self.b = torch.nn.Parameter(torch.randn(5, generator=g))
        self.env_c = torch.nn.Parameter(torch.randn(4, 2, generator=g))

    def forward(self, cam, target=None):
        base = (self.a * cam).sum() + (self.b ** 2).sum() * cam.mean()
        return {"rgb": base * (1.0 + 0.1 * (self.env_c ** 3).sum() * cam.mean())}

    def loss(self, out, target):
        return (out["rgb"] - target) ** 2


def _worker(rank, world, port, q, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from envgs_b200.dist import ViewParallelTrainer, init_process_group_from_env
    init_process_group_from_env("gloo")
    r = _ToyRenderer()
    tr = ViewParallelTrainer(r, average=False, optimizer=False, overlap=overlap)
    cam = torch.full((7, 3), float(rank + 1))
    tr.step(cam, torch.tensor(0.5))
    views = all(tr.flat.grad_is_view(k) for k in range(3))
    q.put((rank, {n: p.grad.tolist() for n, p in r.named_parameters()}, views, tr._early_work is not None, tr.flat.names,
           (tr.flat.late_range, tr.flat.early_range)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_view_parallel_gradient_allreduce_gloo_world2(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + int(overlap)) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    # expected: sum over the two ranks of the single-process gradients
    expect = None
    for rank in range(2):
        r = _ToyRenderer()
        cam = torch.full((7, 3), float(rank + 1))
        r.loss(r(cam), torch.tensor(0.5)).backward()
        g = {n: p.grad.clone() for n, p in r.named_parameters()}
        expect = g if expect is None else {n: expect[n] + g[n] for n in g}
    for rank, grads, is_view, early_launched, names, (late, early) in res:
        assert is_view                                        # parameter grads are views into the single flat buffer
        assert names == ["a", "b", "env_c"] and late[0] == 0 and early[0] == late[1] and late[1] % 8192 == 0
        assert early_launched == overlap                      # the environment group's exchange started from the gradient hook
        for n in expect:
            assert torch.allclose(torch.tensor(grads[n]), expect[n], rtol=1e-5, atol=1e-5), n


def _worker_adam(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from envgs_b200.dist import ViewParallelTrainer, init_process_group_from_env
    init_process_group_from_env("gloo")
    r = _ToyRenderer()
    tr = ViewParallelTrainer(r, lr=1e-2, average=True, optimizer=True)
    loss_host = torch.zeros(())
    losses = []
    for step in range(3):
        cam = torch.full((7, 3), float(rank + 1 + step))
        tr.step(cam, torch.tensor(0.5), loss_out=loss_host)
        losses.append(float(loss_host))
    q.put((rank, {n: p.detach().tolist() for n, p in r.named_parameters()}, losses))
    dist.barrier()
    dist.destroy_process_group()


def test_view_parallel_optimizer_step_gloo_world2():
    """Three optimisation steps on two ranks (gloo, CPU: one all-reduce per group + torch Adam over the flat buffer): the replicas end
    bit-identical and equal a single process that averages the two views' gradients and runs torch.optim.Adam; `loss_out` receives
    each step's loss."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_worker_adam, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    ref = _ToyRenderer()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2, eps=1e-15)
    want_losses = {0: [], 1: []}
    for step in range(3):
        opt.zero_grad()
        for rank in range(2):
            cam = torch.full((7, 3), float(rank + 1 + step))
            loss = ref.loss(ref(cam), torch.tensor(0.5))
            want_losses[rank].append(float(loss.detach()))
            (loss / 2).backward()
        opt.step()
    for n, p in ref.named_parameters():
        assert res[0][1][n] == res[1][1][n], n                       # replicas bit-identical
        assert torch.allclose(torch.tensor(res[0][1][n]), p.detach(), rtol=1e-5, atol=1e-6), n
    for rank in range(2):
        assert res[rank][2] == pytest.approx(want_losses[rank], rel=1e-5)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` without a GPU (the CPU oracle port) prints one JSON line with the keys the driver reads and never
    touches CUDA."""
    import json, subprocess, sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
