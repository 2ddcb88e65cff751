"""Device-side densify / prune / split of a surfel set and the optimizer-state surgery that goes with it.

Host mirror of `GaussianModel.densify_and_prune` and helpers (easyvolcap/utils/gaussian2d_utils.py:517-592, :623-716,
:718-911) over a `GaussianStore`: every parameter of one set lives in ONE packed fp32 buffer `[xyz | f_dc | f_rest | opacity |
scaling | rotation | specular | roughness]`, with gradient and Adam-moment buffers of the same layout (what the view-parallel
trainer all-reduces and the fused Adam kernel walks) plus the four statistics buffers.  A densification step re-packs all of
them with csrc/densify.cu (one plan pass + one scatter launch per array) instead of the reference's per-tensor boolean-mask
gathers, `torch.cat`s and `.item()` syncs; the packed layout — and therefore the all-reduce — stays a single contiguous range.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

PARAM_SPECS = (("_xyz", 3, _lib.ROW_XYZ), ("_features_dc", 3, _lib.ROW_PLAIN), ("_features_rest", 45, _lib.ROW_PLAIN),
               ("_opacity", 1, _lib.ROW_PLAIN), ("_scaling", 2, _lib.ROW_SCALING), ("_rotation", 4, _lib.ROW_PLAIN),
               ("_specular", 1, _lib.ROW_PLAIN), ("_roughness", 1, _lib.ROW_PLAIN))
STAT_SPECS = (("xyz_gradient_accum", _lib.ROW_STAT_GRAD), ("denom", _lib.ROW_STAT_DENOM), ("max_radii2D", _lib.ROW_STAT_RADII),
              ("xyz_weight_accum", _lib.ROW_STAT_WEIGHT))
ROW_FLOATS = sum(w for _, w, _ in PARAM_SPECS)   # 60


class GaussianStore:
    """Packed structure-of-arrays storage of one surfel set: parameters, gradients, Adam moments, statistics."""

    def __init__(self, P: int, device="cuda"):
        dev = torch.device(device)
        self.P, self.device = P, dev
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq = z(ROW_FLOATS * P), z(ROW_FLOATS * P), z(ROW_FLOATS * P), z(ROW_FLOATS * P)
        self.stats = {n: z(P) for n, _ in STAT_SPECS}
        self.step = 0

    @staticmethod
    def from_raw(raw, device="cuda") -> "GaussianStore":
        """From a gaussian_io.RawSurfels (load_ply / load_checkpoint)."""
        st = GaussianStore(raw.P, device)
        for name, w, _ in PARAM_SPECS:
            t = getattr(raw, name)
            if t is None:
                t = torch.logit(torch.full((raw.P, 1), 1e-3 if name == "_specular" else 0.5))
            st.param(name).copy_(t.reshape(raw.P, -1))
        for n, _ in STAT_SPECS:
            if n in raw.buffers:
                st.stats[n].copy_(raw.buffers[n].reshape(-1))
        return st

    def offsets(self, P=None):
        P = self.P if P is None else P
        out, off = {}, 0
        for name, w, _ in PARAM_SPECS:
            out[name] = (off, w)
            off += w * P
        return out

    def _view(self, buf, name):
        off, w = self.offsets()[name]
        return buf[off:off + w * self.P].view(self.P, w)

    def param(self, name):
        return self._view(self.flat, name)

    def grad(self, name):
        return self._view(self.flat_grad, name)

    def moments(self, name):
        return self._view(self.exp_avg, name), self._view(self.exp_avg_sq, name)

    # ---- activations the operators are fed with (gaussian2d_utils.py:362-392) ----
    def activated(self):
        from . import scene
        p = self.param
        shs = torch.cat([p("_features_dc").view(self.P, 1, 3), p("_features_rest").view(self.P, 15, 3)], dim=1)
        return scene.SurfelSet(means=p("_xyz"), scales=torch.exp(p("_scaling")), rotations=torch.nn.functional.normalize(p("_rotation"), dim=-1),
                               opacities=torch.sigmoid(p("_opacity")), shs=shs, specular=torch.sigmoid(p("_specular")),
                               roughness=torch.sigmoid(p("_roughness")))


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def add_densification_stats(store: GaussianStore, viewspace_grad: torch.Tensor, radii: torch.Tensor, weight: torch.Tensor | None = None,
                            update_max_radii: bool = True) -> None:
    """add_densification_stats (:901-911) and the max_radii2D update of envgs_sampler.py:250-253, one kernel; rows with
    radii > 0 are the visibility filter (gaussian2d_utils.py:1143)."""
    g = viewspace_grad.detach().contiguous().float()
    r = radii.contiguous().to(torch.int32)
    w = None if weight is None else weight.detach().contiguous().float()
    s = store.stats
    with torch.cuda.device(store.device):
        _lib.check(_lib.lib().envgs_densify_stats(store.P, g.data_ptr(), r.data_ptr(), None if w is None else w.data_ptr(),
                                                  s["xyz_gradient_accum"].data_ptr(), s["denom"].data_ptr(), s["xyz_weight_accum"].data_ptr(),
                                                  s["max_radii2D"].data_ptr() if update_max_radii else None, _stream(store.device)),
                   "envgs_densify_stats")


def _row_arrays(src: GaussianStore, dst: GaussianStore, with_stats: bool):
    arrs = []
    so, do = src.offsets(), dst.offsets()
    for name, w, role in PARAM_SPECS:
        for sb, db, r in ((src.flat, dst.flat, role), (src.exp_avg, dst.exp_avg, _lib.ROW_OPT_STATE), (src.exp_avg_sq, dst.exp_avg_sq, _lib.ROW_OPT_STATE)):
            arrs.append(_lib.RowArray(sb.data_ptr() + 4 * so[name][0], db.data_ptr() + 4 * do[name][0], w, r))
    if with_stats:
        for n, role in STAT_SPECS:
            arrs.append(_lib.RowArray(src.stats[n].data_ptr(), dst.stats[n].data_ptr(), 1, role))
    return (_lib.RowArray * len(arrs))(*arrs), len(arrs)


def densify_and_prune(store: GaussianStore, min_opacity, min_gradient, densify_grad_threshold, densify_size_threshold, *,
                      split_screen_threshold=None, spatial_scale: float = 1.0, prune_visibility: bool = False, max_gs: int | None = None,
                      max_gs_threshold: float = 0.9, n_split: int = 2, split_ratio: float = 0.8, noise: torch.Tensor | None = None,
                      generator: torch.Generator | None = None) -> GaussianStore:
    """densify_and_prune (:867-899) without `prune_large_gs`: clone -> split -> prune_min_opacity_and_gradients
    [-> prune_visibility] -> reset_stats, returning a NEW store (parameters, Adam moments re-packed; new rows get zero moments
    exactly like cat_tensors_to_optimizer :561-592; gradients zeroed).  `noise`: (n_split, 2, P, 2) standard normals — entry
    [k, 0, i] displaces child k of original i, [k, 1, i] child k of i's clone — drawn here when omitted."""
    L = _lib.lib()
    dev, P = store.device, store.P
    p = _lib.DensifyParams(P, n_split, split_ratio, float(densify_grad_threshold), float(densify_size_threshold * spatial_scale),
                           -1.0 if split_screen_threshold is None else float(split_screen_threshold),
                           -1.0 if min_opacity is None else float(min_opacity), -1.0 if min_gradient is None else float(min_gradient))
    scaling_act = torch.exp(store.param("_scaling")).contiguous()
    opacity_act = torch.sigmoid(store.param("_opacity")).contiguous()
    if noise is None:
        noise = torch.randn(n_split, 2, P, 2, device=dev, generator=generator)
    noise = noise.contiguous().float()
    plan = torch.empty(4 * P, dtype=torch.int32, device=dev)
    counts_dev = torch.zeros(4, dtype=torch.int32, device=dev)
    counts = (C.c_int32 * 4)()
    nbytes = L.envgs_densify_plan_scratch_bytes(P)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = store.stats
    with torch.cuda.device(dev):
        _lib.check(L.envgs_densify_plan(C.byref(p), scaling_act.data_ptr(), opacity_act.data_ptr(), s["xyz_gradient_accum"].data_ptr(),
                                        s["denom"].data_ptr(), s["max_radii2D"].data_ptr(), plan.data_ptr(), counts_dev.data_ptr(), counts,
                                        scratch.data_ptr(), nbytes, _stream(dev)), "envgs_densify_plan")
        n_self, n_clone, n_co, n_cc = (int(c) for c in counts)
        M = n_self + n_clone + n_split * (n_co + n_cc)
        out = GaussianStore(M, dev)
        out.step = store.step
        # running maxima of the weight statistic (densify_stats :693): m0 over the originals, m1 over originals and clones
        w = s["xyz_weight_accum"]
        m0 = w.max() if P else torch.zeros((), device=dev)
        clone_sel = plan[P:2 * P] >= 0
        m1 = torch.maximum(m0, torch.where(clone_sel, w * m0, torch.full_like(w, -float("inf"))).max()) if P else m0
        wmax = torch.stack([m0, m1]).float().contiguous()
        arrs, n = _row_arrays(store, out, with_stats=prune_visibility)
        _lib.check(L.envgs_densify_apply(C.byref(p), plan.data_ptr(), counts, n, arrs, store.param("_xyz").data_ptr(), scaling_act.data_ptr(),
                                         store.param("_rotation").data_ptr(), noise.data_ptr(), wmax.data_ptr(), _stream(dev)),
                   "envgs_densify_apply")
    out.last_counts = dict(kept=n_self, cloned=n_clone, split_originals=n_co, split_clones=n_cc)
    if prune_visibility and max_gs is not None:
        counts_ = out.last_counts
        out = prune_visibility_topk(out, int(max_gs * max_gs_threshold))
        out.last_counts = counts_
    for n_, _ in STAT_SPECS:                                   # reset_stats (:666-674)
        out.stats[n_].zero_()
    return out


def prune_rows(store: GaussianStore, keep_index: torch.Tensor) -> GaussianStore:
    """prune_points + _prune_optimizer + prune_stats (:538-559, :676-687) for an ascending list of surviving rows."""
    idx = keep_index.to(torch.int64).contiguous()
    out = GaussianStore(int(idx.numel()), store.device)
    out.step = store.step
    arrs, n = _row_arrays(store, out, with_stats=True)
    for a in arrs:            # a gather copies rows verbatim whatever their role
        a.role = _lib.ROW_PLAIN
    with torch.cuda.device(store.device):
        _lib.check(_lib.lib().envgs_rows_gather(idx.numel(), idx.data_ptr(), n, arrs, _stream(store.device)), "envgs_rows_gather")
    return out


def prune_visibility_topk(store: GaussianStore, n_after: int) -> GaussianStore:
    """prune_visibility (:849-865): drop the n_before - n_after rows with the smallest average blend weight."""
    n_prune = store.P - n_after
    if n_prune <= 0:
        return store
    avg = store.stats["xyz_weight_accum"] / store.stats["denom"]
    avg = torch.where(avg.isnan(), torch.zeros_like(avg), avg)
    _, idx = torch.topk(avg, n_prune, largest=False)
    keep = torch.ones(store.P, dtype=torch.bool, device=store.device)
    keep[idx] = False
    return prune_rows(store, torch.nonzero(keep)[:, 0])


def reset_opacity(store: GaussianStore, reset_to: float = 0.01) -> None:
    """reset_opacity + replace_tensor_to_optimizer (:512-536): clamp the raw opacity and zero its Adam moments, in place."""
    o = store.param("_opacity")
    o.copy_(torch.minimum(o, torch.logit(torch.full_like(o, reset_to))))
    m, v = store.moments("_opacity")
    m.zero_()
    v.zero_()
