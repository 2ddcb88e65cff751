"""CPU: pin the oracle against golden vectors produced by the UNMODIFIED reference CUDA rasterizer (oracle/make_golden.py,
run on a B200 from oracle/_ref).  If these pass, the oracle restates the reference; the GPU tests then compare the product
kernels with the oracle (and with the same fixtures, tests/test_gpu_golden.py)."""
import os

import pytest
import torch

from oracle import make_golden as MG
from tests import helpers as Hh

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(MG.CASES))
def test_oracle_matches_reference_golden(name):
    gold = torch.load(os.path.join(GOLD, name + ".pt"))
    base, cam, feats, gc, go, bg, use_sh, C = MG.case_inputs(name)
    # the fixtures belong to exactly these seeded inputs
    assert abs(MG.input_checksum(base, cam, feats, gc, go) - gold["input_checksum"]) <= 1e-6 * gold["input_checksum"]
    orc = Hh.run_oracle_raster(base, cam, C, feats=feats, use_sh=use_sh, g_color=gc, g_others=go, bg=bg)
    assert orc["num_rendered"] == gold["num_rendered"]          # same tile instancing (getRect semantics)
    assert torch.equal(orc["radii"], gold["radii"])
    for key in ("color", "allmap"):
        scale = max(1.0, float(gold[key].abs().max())) if key == "allmap" else 1.0
        assert Hh.frac_above(orc[key], gold[key], 1e-4 * scale) <= 1e-3, Hh.image_report(key, orc[key], gold[key])
        assert float((orc[key] - gold[key]).abs().max()) <= 3e-2 * scale, Hh.image_report(key, orc[key], gold[key])
    assert Hh.rel_l2(orc["weight"], gold["weight"]) < 1e-4
    for k in [k for k in gold if k.startswith("g_")]:
        t, full = Hh.trimmed_rel_l2(orc[k], gold[k]), Hh.rel_l2(orc[k], gold[k])
        # the densification signal dL/dT[.,2] * depth * W/2 (backward.cu:626-629) is a sum of cancelling per-pixel terms:
        # fp32 (reference) and fp64 (oracle) agree to ~1e-2 only; every other gradient must meet 1e-3
        tol = 2e-2 if k == "g_means2D" else 1e-3
        assert t <= tol, f"{k}: trimmed rel L2 {t:.3e} (full {full:.3e})"
        assert full <= 5e-2, f"{k}: rel L2 {full:.3e}"


def test_reference_atomics_noise_is_below_tolerance():
    for name in MG.CASES:
        gold = torch.load(os.path.join(GOLD, name + ".pt"))
        assert gold["noise_g_means"] < 1e-5
