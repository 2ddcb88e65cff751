"""GPU: device-side densify / prune / split (csrc/densify.cu through envgs_b200.densify) against the PyTorch twin of the
reference code (tests/densify_twin.py) on seeded inputs: same surviving rows in the same order, same parameters, same Adam
moments (zeros for new rows), bit-exact wherever the arithmetic is a copy."""
import pytest
import torch

from envgs_b200 import densify as D
from envgs_b200 import gaussian_io as IO
from envgs_b200 import scene
from tests.densify_twin import NAMES, TwinModel

pytestmark = pytest.mark.gpu


def _setup(P, seed):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw = IO.RawSurfels.from_activated(scene.make_base_set(P, seed=seed, s0=0.02))
    store = D.GaussianStore.from_raw(raw, dev)
    # a few optimisation steps' worth of state: moments and statistics with structure
    store.exp_avg.copy_(torch.randn(store.exp_avg.shape, generator=g) * 1e-3)
    store.exp_avg_sq.copy_(torch.rand(store.exp_avg_sq.shape, generator=g) * 1e-6)
    denom = torch.randint(0, 6, (P,), generator=g).float()
    store.stats["denom"].copy_(denom)
    store.stats["xyz_gradient_accum"].copy_(torch.rand(P, generator=g) * 4e-4 * denom)
    store.stats["max_radii2D"].copy_(torch.rand(P, generator=g) * 40)
    store.stats["xyz_weight_accum"].copy_(torch.rand(P, generator=g) * 3 * denom)
    noise = torch.randn(2, 2, P, 2, generator=g).to(dev)
    return store, noise


def _twin_of(store, noise):
    params = {n: store.param(n).clone().view(store.P, *([1, 3] if n == "_features_dc" else [15, 3] if n == "_features_rest" else [-1]))
              for n in NAMES}
    tw = TwinModel(params, {k: v.clone() for k, v in store.stats.items()}, noise)
    for n in NAMES:                     # hand the twin's Adam the same moments
        p = getattr(tw, n)
        m, v = store.moments(n)
        tw.optimizer.state[p] = {"step": torch.tensor(5.0), "exp_avg": m.clone().view_as(p), "exp_avg_sq": v.clone().view_as(p)}
    return tw


@pytest.mark.parametrize("P,seed,screen,prune_vis", [(5000, 0, None, False), (20000, 1, 20.0, False), (12000, 2, 20.0, True)])
def test_densify_and_prune_matches_reference_twin(built_lib, P, seed, screen, prune_vis):
    store, noise = _setup(P, seed)
    tw = _twin_of(store, noise)
    kw = dict(min_opacity=0.05, min_gradient=2e-5, grad_thr=2e-4, size_thr=0.02)
    n_after = None
    tw.densify_and_prune(kw["min_opacity"], kw["min_gradient"], kw["grad_thr"], kw["size_thr"], split_screen_threshold=screen)
    if prune_vis:
        n_after = int(tw._xyz.shape[0] * 0.9)
        tw2 = _twin_of(store, noise)
        tw2.densify_and_prune(kw["min_opacity"], kw["min_gradient"], kw["grad_thr"], kw["size_thr"], split_screen_threshold=screen,
                              prune_visibility=True, n_after=n_after)
        tw = tw2
    out = D.densify_and_prune(store, kw["min_opacity"], kw["min_gradient"], kw["grad_thr"], kw["size_thr"], split_screen_threshold=screen,
                              prune_visibility=prune_vis, max_gs=n_after, max_gs_threshold=1.0, noise=noise)
    c = out.last_counts
    assert c["cloned"] > 0 and c["split_originals"] > 0 and out.P != P, c           # the case exercises every branch
    if screen is not None:
        assert c["split_clones"] > 0, c
    assert out.P == tw._xyz.shape[0], (out.P, tw._xyz.shape[0], c)
    for n in NAMES:
        a, b = out.param(n), getattr(tw, n).detach().reshape(out.P, -1)
        if n == "_xyz":       # children positions go through bmm in the twin: same formula, different FMA order
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), n
            same = (a == b).all(dim=1)
            assert int(same.sum()) >= c["kept"]                                     # copied rows are bit-exact
        else:
            assert torch.equal(a, b), n
        m, v = out.moments(n)
        tm, tv = tw.moments(n)
        assert torch.equal(m, tm.reshape(out.P, -1)) and torch.equal(v, tv.reshape(out.P, -1)), n
    for k in out.stats:                                                             # reset_stats
        assert float(out.stats[k].abs().max()) == 0.0
    assert float(out.flat_grad.abs().max()) == 0.0


def test_stats_kernel_matches_reference_update(built_lib):
    P = 4000
    store, _ = _setup(P, 3)
    g = torch.Generator(device="cpu").manual_seed(9)
    vs = (torch.randn(P, 3, generator=g) * 1e-4).cuda()
    radii = torch.randint(-1, 30, (P,), generator=g).int().cuda()
    weight = torch.rand(P, 1, generator=g).cuda()
    ref = {k: v.clone() for k, v in store.stats.items()}
    vis = radii > 0
    ref["denom"][vis] += 1
    ref["xyz_gradient_accum"][vis] += torch.norm(vs[vis], dim=-1)
    ref["xyz_weight_accum"][vis] += weight[vis, 0]
    ref["max_radii2D"][vis] = torch.max(ref["max_radii2D"][vis], radii[vis].float())
    D.add_densification_stats(store, vs, radii, weight)
    for k in ref:
        assert torch.allclose(store.stats[k], ref[k], rtol=1e-6, atol=0), k


def test_prune_rows_and_reset_opacity(built_lib):
    store, _ = _setup(3000, 4)
    keep = torch.nonzero(torch.rand(3000, device="cuda") > 0.3)[:, 0]
    out = D.prune_rows(store, keep)
    assert out.P == keep.numel()
    for n in NAMES:
        assert torch.equal(out.param(n), store.param(n)[keep])
        assert torch.equal(out.moments(n)[0], store.moments(n)[0][keep])
    assert torch.equal(out.stats["denom"], store.stats["denom"][keep])
    D.reset_opacity(out, 0.01)
    assert float(torch.sigmoid(out.param("_opacity")).max()) <= 0.01 + 1e-6
    assert float(out.moments("_opacity")[0].abs().max()) == 0.0
