"""Golden vectors for densify / prune / split from the UNMODIFIED reference code, generated in the build container (CPU).

Test infrastructure only.  Runs `GaussianModel.densify_and_prune` of /root/reference/easyvolcap/utils/gaussian2d_utils.py
(:867-899 and everything it calls: clone/split :718-760, optimizer surgery :517-592, statistics :623-716, prune rules :762-865)
on a seeded surfel set with a real `torch.optim.Adam` holding non-trivial moments, and stores inputs + outputs in
tests/golden/densify_ref_*.pt.  tests/test_oracle_golden.py checks tests/densify_twin.py (the restatement the device
implementation is compared with on the GPU, tests/test_gpu_densify.py) against these files, which closes the chain
reference -> twin -> CUDA.

Only gaussian2d_utils.py itself is executed.  Its helper imports are replaced by the few stubs below: the real
`easyvolcap.utils.console_utils` pulls in the whole EasyVolcap tree and tries to install missing packages at import time, `plyfile`
is not in this image, and `net_utils.make_params / make_buffer` are two one-liners (net_utils.py:172-177) restated here.

The one source of randomness, `torch.normal(means, stds)` in `split` (:746), is replaced for the duration of the call by
`means + stds * z` with z taken from a seeded table indexed by (copy, row is a clone, original row) — the same table the twin and the
device implementation consume — so the three can be compared value by value.  The provenance (original row, clone flag) of every
current row is tracked from outside by wrapping `clone`, `split` and `prune_points`; the wrapped methods themselves run unmodified.

    python oracle/make_densify_golden.py        # writes tests/golden/densify_ref_*.pt
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_FILE = "/root/reference/easyvolcap/utils/gaussian2d_utils.py"
NAMES = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_specular", "_roughness")

CASES = {
    # name: (P, seed, split_screen_threshold, prune_visibility)
    "densify_ref_basic": (600, 0, None, False),
    "densify_ref_screen": (800, 1, 20.0, False),
    "densify_ref_visibility": (700, 2, 20.0, True),
}
RULES = dict(min_opacity=0.05, min_gradient=2e-5, densify_grad_threshold=2e-4, densify_size_threshold=0.02)


def load_reference_module(ref_file: str = REF_FILE, mod_name: str = "ref_gaussian2d_utils", extra_stubs: dict | None = None):
    class dotdict(dict):
        __getattr__ = dict.get
        __setattr__ = dict.__setitem__

    ident = lambda s, *a, **k: s
    console = types.ModuleType("easyvolcap.utils.console_utils")
    console.dotdict = dotdict
    console.log = lambda *a, **k: None
    for name in ("magenta", "yellow_slim", "yellow", "red", "blue", "green", "cyan"):
        setattr(console, name, ident)
    net = types.ModuleType("easyvolcap.utils.net_utils")
    net.make_params = lambda p: nn.Parameter(torch.as_tensor(p), requires_grad=True)       # net_utils.py:172-173
    net.make_buffer = lambda p: nn.Parameter(torch.as_tensor(p), requires_grad=False)      # net_utils.py:176-177
    net.typed = lambda *a, **k: (lambda f: f)
    sh = types.ModuleType("easyvolcap.utils.sh_utils")
    sh.eval_sh = None
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = object
    pkg, utils = types.ModuleType("easyvolcap"), types.ModuleType("easyvolcap.utils")
    pkg.__path__, utils.__path__ = [], []
    stubs = {"easyvolcap": pkg, "easyvolcap.utils": utils, "easyvolcap.utils.console_utils": console,
             "easyvolcap.utils.net_utils": net, "easyvolcap.utils.sh_utils": sh, "plyfile": ply}
    stubs.update(extra_stubs or {})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(mod_name, ref_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, dotdict


def case_inputs(name):
    """Seeded inputs, the same recipe as tests/test_gpu_densify.py::_setup (CPU tensors)."""
    from envgs_b200 import gaussian_io as IO
    from envgs_b200 import scene
    P, seed, screen, prune_vis = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw = IO.RawSurfels.from_activated(scene.make_base_set(P, seed=seed, s0=0.02))
    params = {n: getattr(raw, n).clone().float() for n in NAMES}
    moments = {n: (torch.randn(params[n].shape, generator=g) * 1e-3, torch.rand(params[n].shape, generator=g) * 1e-6) for n in NAMES}
    denom = torch.randint(0, 6, (P,), generator=g).float()
    stats = dict(denom=denom, xyz_gradient_accum=torch.rand(P, generator=g) * 4e-4 * denom, max_radii2D=torch.rand(P, generator=g) * 40,
                 xyz_weight_accum=torch.rand(P, generator=g) * 3 * denom)
    noise = torch.randn(2, 2, P, 2, generator=g)
    return params, moments, stats, noise, screen, prune_vis


def run_reference(name):
    G, dotdict = load_reference_module()
    params, moments, stats, noise, screen, prune_vis = case_inputs(name)
    P = params["_xyz"].shape[0]
    model = G.GaussianModel(xyz=params["_xyz"].clone(), colors=None, init_scale=params["_scaling"].clone(), xyz_lr_scheduler=None,
                            render_reflection=True)
    for n in NAMES:                                                   # exactly the seeded (pre-activation) values
        setattr(model, n, nn.Parameter(params[n].clone(), requires_grad=True))
    model.max_radii2D.data = stats["max_radii2D"].clone()
    model.xyz_gradient_accum.data = stats["xyz_gradient_accum"].clone()[:, None]
    model.denom.data = stats["denom"].clone()[:, None]
    model.xyz_weight_accum.data = stats["xyz_weight_accum"].clone()[:, None]
    opt = torch.optim.Adam([{"params": [getattr(model, n)], "name": n} for n in NAMES], lr=1e-3, eps=1e-15)
    for n in NAMES:
        opt.state[getattr(model, n)] = {"step": torch.tensor(5.0), "exp_avg": moments[n][0].clone(), "exp_avg_sq": moments[n][1].clone()}

    # ---- provenance tracking + the noise table in place of torch.normal ----
    prov = dict(src=torch.arange(P), is_clone=torch.zeros(P, dtype=torch.long), pending=None)
    cls = G.GaussianModel
    orig_clone, orig_split, orig_prune = cls.clone, cls.split, cls.prune_points

    def clone(self, mask, *a, **k):
        orig_clone(self, mask, *a, **k)
        prov["src"] = torch.cat([prov["src"], prov["src"][mask]])
        prov["is_clone"] = torch.cat([prov["is_clone"], torch.ones_like(prov["is_clone"][mask])])

    def split(self, mask, N=2, *a, **k):
        src, isc = prov["src"][mask], prov["is_clone"][mask]
        prov["pending"] = (src, isc, N)
        prov["src"] = torch.cat([prov["src"], src.repeat(N)])          # appended rows; the originals go in the prune_points call inside
        prov["is_clone"] = torch.cat([prov["is_clone"], isc.repeat(N)])
        orig_split(self, mask, N, *a, **k)
        prov["pending"] = None

    def prune_points(self, mask, *a, **k):
        orig_prune(self, mask, *a, **k)
        prov["src"], prov["is_clone"] = prov["src"][~mask], prov["is_clone"][~mask]

    def table_normal(means, stds):
        src, isc, N = prov["pending"]
        z = torch.cat([noise[k, isc, src] for k in range(N)], dim=0)
        return means + stds * torch.cat([z, torch.zeros_like(z[:, :1])], dim=-1)

    snapshot = {}
    orig_reset = cls.reset_stats

    def reset_stats(self):                                            # densify_and_prune ends with it: keep the statistics it clears
        snapshot.update(xyz_gradient_accum=self.xyz_gradient_accum.detach().clone()[:, 0], denom=self.denom.detach().clone()[:, 0],
                        max_radii2D=self.max_radii2D.detach().clone(), xyz_weight_accum=self.xyz_weight_accum.detach().clone()[:, 0])
        orig_reset(self)

    cls.clone, cls.split, cls.prune_points, cls.reset_stats = clone, split, prune_points, reset_stats
    real_normal, real_empty_cache = torch.normal, torch.cuda.empty_cache
    torch.normal, torch.cuda.empty_cache = table_normal, (lambda: None)
    n_after = None
    try:
        if prune_vis:     # the count to keep: 90 % of what the same call leaves without the visibility pruning (as the GPU test does)
            n_after = int(run_reference_count(name) * 0.9)
            model.max_gs, model.max_gs_threshold = n_after, 1.0
        with torch.no_grad():
            model.densify_and_prune(optimizer=opt, split_screen_threshold=screen, prune_visibility=prune_vis, **RULES)
    finally:
        torch.normal, torch.cuda.empty_cache = real_normal, real_empty_cache
        cls.clone, cls.split, cls.prune_points, cls.reset_stats = orig_clone, orig_split, orig_prune, orig_reset
    out = {"params": {n: getattr(model, n).detach().clone() for n in NAMES},
           "exp_avg": {n: opt.state[getattr(model, n)]["exp_avg"].clone() for n in NAMES},
           "exp_avg_sq": {n: opt.state[getattr(model, n)]["exp_avg_sq"].clone() for n in NAMES},
           "stats_before_reset": snapshot, "src": prov["src"].clone(), "is_clone": prov["is_clone"].clone(),
           "stats_after": {"max_radii2D": model.max_radii2D.detach().clone(), "denom": model.denom.detach().clone()}}
    # the inputs are a pure function of the case (case_inputs): only their checksum travels with the outputs
    return {"rules": dict(RULES, split_screen_threshold=screen, prune_visibility=prune_vis, n_after=n_after),
            "input_checksum": input_checksum(params, moments, stats, noise), "outputs": out,
            "reference": "zju3dv/EnvGS easyvolcap/utils/gaussian2d_utils.py GaussianModel.densify_and_prune, unmodified, CPU, torch " + torch.__version__}


def input_checksum(params, moments, stats, noise) -> float:
    parts = list(params.values()) + [m for pair in moments.values() for m in pair] + list(stats.values()) + [noise]
    return float(sum(p.double().abs().sum() for p in parts))


_COUNT_CACHE = {}


def run_reference_count(name):
    """Row count the reference leaves for this case WITHOUT visibility pruning (defines n_after of the visibility case)."""
    if name not in _COUNT_CACHE:
        P, seed, screen, _ = CASES[name]
        CASES["_tmp"] = (P, seed, screen, False)
        try:
            _COUNT_CACHE[name] = run_reference("_tmp")["outputs"]["params"]["_xyz"].shape[0]
        finally:
            CASES.pop("_tmp")
    return _COUNT_CACHE[name]


if __name__ == "__main__":
    outdir = os.path.join(ROOT, "tests", "golden")
    for name in list(CASES):
        res = run_reference(name)
        torch.save(res, os.path.join(outdir, name + ".pt"))
        o = res["outputs"]
        print(name, "rows", CASES[name][0], "->", o["params"]["_xyz"].shape[0], "clones kept", int(o["is_clone"].sum()), "n_after",
              res["rules"]["n_after"])
