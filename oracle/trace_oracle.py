"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the reference surfel tracer.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module.  The product path (envgs_b200/*) never imports it.

Brute force: every ray is intersected with every surfel quad, hits are sorted by distance and composited with the
exact rules of
  submodules/diff-surfel-tracing/optix_tracer/
      forward.cu:219-450   traceRay      (per-hit evaluation, thresholds, 256-contributor cap)
      forward.cu:453-559   tracePath     (segments, specular chain, reflection)
      forward.cu:77-117    compute_transmat_uv
      forward.cu:120-173   compute_transmat_xy (first segment when start_from_first)
      forward.cu:29-72     SH colour along the ray direction
  easyvolcap/utils/optix_utils.py:39-69  (quad = +-3 sigma square, one hit per surfel)
Backward comes from autograd, with the reference's deviations reproduced (backward.cu):
  * alpha cap straight-through (`power_clamped = 1`, :801-802)
  * quaternion normalisation not back-propagated (auxiliary.h:444-488)
  * the hit distance is differentiated as t = m.(mu-o)/(m.d), m = t_u x t_v — algebraically the reference's
    quad-vertex form (:293-306, :319-334) — while the accumulated normal uses R[:,2]
  * no gradient for `dist` (:796) and, unless start_from_first, none for `acc` (:792-796)
  * non-zero background would be counted twice in dL/dalpha by the reference (:769-772); use bg = 0 when comparing
  * two segments: a ray whose second segment is skipped back-propagates its first colour scaled by spec_0 (:955-957, :988)
Third-party arithmetic outside /root/reference: the ray/triangle distance `t` comes from NVIDIA OptiX 7.7 (closed
source, driver-side libnvoptix; headers vendored at third_party/optix @ 7b5c4e86).  Its specification here is the
analytic ray-plane distance the reference's own backward differentiates (backward.cu:293-302).
Pinning status: PINNED by tests/golden/trace_ref_*.pt — outputs of the UNMODIFIED reference device programs
(optix_tracer/forward.cu, backward.cu) compiled against the OptiX emulation header oracle/optix_shim/optix.h (recipe:
oracle/build_ref.py, generator: oracle/make_golden.py, checked by tests/test_oracle_golden.py).  What remains outside the
pin is only OptiX's own ray/triangle distance, emulated there by fp64 Moeller-Trumbore.
Chunk restart (forward.cu:288-441): the reference fetches hits 16 at a time and restarts the ray at
t_last + STEP_EPSILON, so a hit within 1e-5 (in t) behind every 16th hit is never seen; `_chunk_restart_keep` restates it.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .raster_oracle import FAR_N, FILTER_INV_SQUARE, NEAR_N, eval_sh_basis, quat_to_rotmat

START_OFFSET = 0.06
STEP_EPSILON = 1e-5
MAX_INTERSECTION = 256
MID = dict(rayo=0, rayd=3, dpt=6, acc=7, norm=8, aux=11, rgb=13)


@dataclass
class TraceOut:
    rgb: torch.Tensor    # (H,W,3)
    dpt: torch.Tensor    # (H,W,1)
    acc: torch.Tensor    # (H,W,1)
    norm: torch.Tensor   # (H,W,3)
    dist: torch.Tensor   # (H,W,1)
    aux: torch.Tensor    # (H,W,2)
    mid: torch.Tensor    # (H,W,32)
    wet: torch.Tensor    # (P,1)
    border: torch.Tensor = None   # (H,W) bool: a discrete decision on this ray lies within fp32 noise of its threshold


# Relative width of the "either side is right" band around the discrete per-hit decisions (alpha >= 1/255, |u|,|v| <= 3,
# cmd >= min depth, rho3d <= rho2d).  fp32 evaluation of a hit at distance t ~ 20 carries ~5 ulp of position error, i.e.
# ~7e-5 relative in alpha (tools/trace_diag.py, profiles/round2_parity_experiments.md): the reference's own fp32 result
# and any other fp32 evaluation (including real OptiX's hit distance vs the emulation's) may land on different sides.
BORDER_REL = 3e-4
BORDER_REL_LATER = 1.5e-3   # later segments start from the previous segment's fp32 depth/normal: their rays carry its noise too


def _chunk_restart_keep(tk: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """Which entries of the per-ray sorted hit list (r,k) the reference's chunked traversal ever sees (forward.cu:288-441):
    after a full chunk of CHUNK_SIZE = 16 hits the next optixTrace starts at ray_o + (t_16th + STEP_EPSILON) * ray_d with
    tmin = 0, so hits with t <= t_16th + 1e-5 beyond the 16th are dropped."""
    r, k = tk.shape
    keep = hk.clone()
    nh = hk.sum(1)
    start = torch.zeros(r, dtype=torch.long, device=tk.device)
    idx = torch.arange(k, device=tk.device)[None]
    for _ in range((k + 15) // 16):
        e = start + 15
        valid = e < nh
        if not bool(valid.any()):
            break
        t_end = tk.gather(1, e.clamp(max=k - 1)[:, None])[:, 0]
        drop = valid[:, None] & (idx > e[:, None]) & hk & (tk <= (t_end + STEP_EPSILON)[:, None])
        keep &= ~drop
        start = torch.where(valid, e + 1 + drop.sum(1), torch.full_like(start, k))
    return keep


def _segment(ray_o, ray_d, pix, active, means, scales, rotations, opacities, shs, sh_degree, colors, others, campos, bg,
             first, projmatrix, W, H, scale_modifier, max_hits, ray_chunk, border_rel=BORDER_REL):
    """Trace one segment for R rays.  Returns C(R,3), D, A, N(R,3), dist, M1, M2, O(R,2), wet(P)."""
    dt = means.dtype
    dev = means.device
    R_ = ray_o.shape[0]
    P = means.shape[0]
    Rm = quat_to_rotmat(rotations)
    tu, tv, nr = Rm[:, :, 0], Rm[:, :, 1], Rm[:, :, 2]
    mvec = torch.linalg.cross(tu, tv, dim=-1)
    dirc = means - campos[None]
    cosv = -(dirc * nr).sum(-1)
    nsign = torch.where(cosv > 0, 1.0, -1.0).to(dt)
    nflip = nr * nsign[:, None]
    min_depth = NEAR_N if first else START_OFFSET

    outs = []
    wet = torch.zeros(P, dtype=dt, device=dev)
    for s in range(0, R_, ray_chunk):
        o, d = ray_o[s:s + ray_chunk], ray_d[s:s + ray_chunk]
        act = active[s:s + ray_chunk]
        r = o.shape[0]
        with torch.no_grad():   # candidate search: ray/plane distance + +-3 sigma square test
            q = d @ mvec.t()                                        # (r,P)
            num = ((means[None] - o[:, None]) * mvec[None]).sum(-1)
            qs = torch.where(q != 0, q, torch.ones_like(q))
            t = num / qs
            x = o[:, None] + t[..., None] * d[:, None]
            u = ((x * tu[None]).sum(-1) - (means * tu).sum(-1)[None]) / scales[None, :, 0]
            v = ((x * tv[None]).sum(-1) - (means * tv).sum(-1)[None]) / scales[None, :, 1]
            hit = (q != 0) & (t > 0) & (t < 1e16) & (u.abs() <= 3) & (v.abs() <= 3) & act[:, None] & torch.isfinite(t)
            emax = torch.maximum(u.abs(), v.abs())
            near_edge = ((q != 0) & (t > 0) & act[:, None] & torch.isfinite(t) & ((emax - 3.0).abs() <= 3.0 * border_rel)).any(1)
            tkey = torch.where(hit, t, torch.full_like(t, float("inf")))
            k = min(max_hits, P)
            tk, ik = torch.topk(tkey, k, dim=1, largest=False, sorted=True)       # (r,k)
            hk = torch.isfinite(tk)
            if k < P and bool((hit.sum(1) > k).any()):
                raise RuntimeError("trace oracle: max_hits too small for this scene")
            hk = _chunk_restart_keep(tk, hk)
        # differentiable re-evaluation of the selected hits
        mu, sc, mv = means[ik], scales[ik], mvec[ik]                  # (r,k,3) ...
        qd = (mv * d[:, None]).sum(-1)
        qd = torch.where(hk, qd, torch.ones_like(qd))
        tt = ((mu - o[:, None]) * mv).sum(-1) / qd
        tt = torch.where(hk, tt, torch.zeros_like(tt))
        xyz = o[:, None] + tt[..., None] * d[:, None]
        tuk, tvk = tu[ik], tv[ik]
        uu = ((tuk * xyz).sum(-1) - (tuk * mu).sum(-1)) / sc[..., 0]
        vv = ((tvk * xyz).sum(-1) - (tvk * mu).sum(-1)) / sc[..., 1]
        rho3d = uu * uu + vv * vv
        normal = nflip[ik]
        cos_ok = (cosv[ik] != 0)
        if first:
            PV = projmatrix.to(dt).t()
            N = torch.tensor([[W / 2.0, 0, 0, (W - 1) / 2.0], [0, H / 2.0, 0, (H - 1) / 2.0], [0, 0, 0, 1.0]], dtype=dt, device=dev)
            L0 = tu * (scale_modifier * scales[:, 0:1])
            L1 = tv * (scale_modifier * scales[:, 1:2])
            zeros, ones = torch.zeros(P, 1, dtype=dt, device=dev), torch.ones(P, 1, dtype=dt, device=dev)
            S = torch.stack([torch.cat([L0, zeros], -1), torch.cat([L1, zeros], -1), torch.cat([means, ones], -1)], dim=-1)
            s2p = (N @ PV) @ S                                            # (P,3,3)
            tvec = torch.tensor([9.0, 9.0, -1.0], dtype=dt, device=dev)
            dd = (tvec * s2p[:, 2] * s2p[:, 2]).sum(-1)
            okxy = dd != 0
            f = tvec[None] / torch.where(okxy, dd, torch.ones_like(dd))[:, None]
            cxy = torch.stack([(f * s2p[:, 0] * s2p[:, 2]).sum(-1), (f * s2p[:, 1] * s2p[:, 2]).sum(-1)], -1)
            px = pix[s:s + ray_chunk]                                     # (r,2) = (col,row)
            dxy = cxy[ik] - px[:, None].to(dt)
            rho2d = (dxy * dxy).sum(-1) * FILTER_INV_SQUARE
            skip_xy = (~okxy[ik]) & (rho3d > rho2d)
            cmd = torch.where(rho3d <= rho2d, tt, s2p[ik][:, :, 2, 2])
        else:
            rho2d = rho3d + 1.0
            skip_xy = torch.zeros_like(hk)
            cmd = tt
        power = -0.5 * torch.minimum(rho3d, rho2d)
        oG = opacities[ik, 0] * torch.exp(power)
        alpha = oG + (torch.clamp(oG, max=0.99) - oG).detach()
        cand = hk & cos_ok & (~skip_xy) & (cmd >= min_depth) & (power <= 0) & (alpha >= 1.0 / 255.0)
        with torch.no_grad():
            thr = 1.0 / 255.0
            nb_ = hk & cos_ok & (((alpha - thr).abs() <= border_rel * thr) | ((cmd - min_depth).abs() <= border_rel * min_depth))
            if first:
                nb_ = nb_ | (hk & ((rho3d - rho2d).abs() <= border_rel * rho2d.abs()))
        one_m = torch.where(cand, 1.0 - alpha, torch.ones_like(alpha))
        Tincl = torch.cumprod(one_m, dim=1)
        Tbefore = torch.cat([torch.ones(r, 1, dtype=dt, device=dev), Tincl[:, :-1]], dim=1)
        stop = cand & (Tincl.detach() < 0.0001)
        alive = torch.cumsum(stop.to(torch.int32), dim=1) == 0
        nth = torch.cumsum((cand & alive).to(torch.int32), dim=1)          # 1-based contributor count
        contrib = cand & alive & (nth <= MAX_INTERSECTION)
        with torch.no_grad():   # hits behind the termination point cannot change the result
            alive_before = torch.cat([torch.ones(r, 1, dtype=torch.bool, device=dev), alive[:, :-1]], dim=1)
            border = near_edge | (nb_ & alive_before).any(1) | (cand & ((Tincl - 0.0001).abs() <= 1e-2 * 0.0001) & alive_before).any(1)
        w = torch.where(contrib, alpha * Tbefore, torch.zeros_like(alpha))
        Tfinal = torch.where(contrib, one_m, torch.ones_like(one_m)).prod(dim=1)

        if colors is None:
            nb = (sh_degree + 1) ** 2
            dn = d / d.norm(dim=-1, keepdim=True)
            basis = eval_sh_basis(sh_degree, dn)                          # (r,nb)
            col = torch.clamp_min((basis[:, None, :, None] * shs[ik][:, :, :nb, :]).sum(2) + 0.5, 0.0)   # (r,k,3)
        else:
            col = colors[ik]
        Cc = (w[..., None] * col).sum(1) + Tfinal[:, None] * bg[None]
        if others is not None:
            O = (w[..., None] * others[ik]).sum(1)
        else:
            O = torch.zeros(r, 2, dtype=dt, device=dev)
        csafe = torch.where(contrib, cmd, torch.ones_like(cmd))
        m = FAR_N / (FAR_N - NEAR_N) * (1.0 - NEAR_N / csafe)
        mw, mmw = m * w, m * m * w
        M1_ex, M2_ex = torch.cumsum(mw, 1) - mw, torch.cumsum(mmw, 1) - mmw
        dist = ((m * m * (1.0 - Tbefore) + M2_ex - 2.0 * m * M1_ex) * w).sum(1)
        D = (w * csafe).sum(1)
        Nn = (w[..., None] * normal).sum(1)
        A = 1.0 - Tfinal
        wet = wet.index_add(0, ik.reshape(-1), w.detach().reshape(-1))
        outs.append((Cc, D, A, Nn, dist, mw.sum(1), mmw.sum(1), O, border))
    cat = lambda i: torch.cat([o[i] for o in outs], dim=0)
    return cat(0), cat(1), cat(2), cat(3), cat(4), cat(5), cat(6), cat(7), wet, cat(8)


def trace(ray_o, ray_d, means3D, opacities, scales, rotations, shs=None, colors_precomp=None, others_precomp=None, *,
          sh_degree=3, campos, bg, start_from_first=False, max_trace_depth=0, specular_threshold=0.0, projmatrix=None,
          scale_modifier=1.0, training=True, max_hits=96, ray_chunk=2048) -> TraceOut:
    dt = means3D.dtype
    dev = means3D.device
    H, W = ray_o.shape[0], ray_o.shape[1]
    R_ = H * W
    o = ray_o.reshape(R_, 3)
    d = ray_d.reshape(R_, 3)
    rows, cols = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    pix = torch.stack([cols.reshape(-1), rows.reshape(-1)], -1)
    campos = campos.to(dt)
    bgv = bg.reshape(-1)[:3].to(dt)

    rgb = torch.zeros(R_, 3, dtype=dt, device=dev)
    mid = torch.zeros(R_, 32, dtype=dt, device=dev)
    s_prod = torch.ones(R_, dtype=dt, device=dev)
    active = torch.ones(R_, dtype=torch.bool, device=dev)
    wet_total = torch.zeros(means3D.shape[0], dtype=dt, device=dev)
    out_dpt = out_acc = out_norm = out_dist = out_aux = None
    mids = []
    for i in range(max_trace_depth + 1):
        first = (i == 0) and start_from_first
        Cc, D, A, Nn, dist, M1, M2, O, wet, brd = _segment(o, d, pix, active, means3D, scales, rotations, opacities, shs, sh_degree,
                                                      colors_precomp, others_precomp, campos, bgv, first, projmatrix, W, H,
                                                      scale_modifier, max_hits, ray_chunk,
                                                      BORDER_REL if i == 0 else BORDER_REL_LATER)
        wet_total = wet_total + wet
        border_total = brd if i == 0 else (border_total | (brd & active))
        a = active.to(dt)
        if i == 0:
            out_dpt, out_norm, out_aux = D, Nn, O
            out_acc = A if start_from_first else A.detach()
            out_dist = dist.detach()
        f_this = torch.ones_like(s_prod) if i == max_trace_depth else 1.0 - O[:, 0]
        nlen = Nn.norm(dim=-1)
        cont = active & (nlen != 0) & (O[:, 0] > specular_threshold)
        Cc_bp = Cc
        if i < max_trace_depth:
            # Reference deviation (backward.cu:955-957, :988): the running specular product is divided by spec_i only when the
            # NEXT segment is actually walked, so a ray that stops here back-propagates its colour scaled by spec_i.  Value
            # unchanged, gradient of the colour scaled.
            q = torch.where(active & ~cont, O[:, 0].detach(), torch.ones_like(s_prod))
            Cc_bp = Cc.detach() + (Cc - Cc.detach()) * q[:, None]
        rgb = rgb + a[:, None] * Cc_bp * (f_this * s_prod)[:, None]
        s_prod = s_prod * torch.where(active, O[:, 0], torch.ones_like(s_prod))
        seg = torch.cat([o, d, D[:, None], A[:, None], Nn, O, Cc], dim=-1) * a[:, None]
        mids.append(seg)
        if i == max_trace_depth:
            break
        nsafe = torch.where(cont[:, None], Nn, torch.ones_like(Nn))
        n = nsafe / nsafe.norm(dim=-1, keepdim=True)
        d_new = d - 2.0 * (n * d).sum(-1, keepdim=True) * n
        E = o + D[:, None] * d
        o_new = E + STEP_EPSILON * d_new
        o = torch.where(cont[:, None], o_new, o)
        d = torch.where(cont[:, None], d_new, d)
        active = cont
    while len(mids) < 2:
        mids.append(torch.zeros(R_, 16, dtype=dt, device=dev))
    mid = torch.cat(mids, dim=-1).detach()
    return TraceOut(rgb.reshape(H, W, 3), out_dpt.reshape(H, W, 1), out_acc.reshape(H, W, 1), out_norm.reshape(H, W, 3),
                    out_dist.reshape(H, W, 1), out_aux.reshape(H, W, 2), mid.reshape(H, W, 32), wet_total[:, None],
                    border_total.reshape(H, W))
