"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the reference surfel rasterizer.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / CPU baseline.  The product path (envgs_b200/*) never imports it.

Restates, statement by statement, the algorithm of
  submodules/diff-surfel-rasterizations/diff-surfel-rasterization-wet-ch05/cuda_rasterizer/
      forward.cu:75-115   compute_transmat        (homography T = rows of N * PV * S)
      forward.cu:119-145  compute_aabb            (projected centre + 3-sigma extent)
      forward.cu:148-251  preprocessCUDA          (near cull, normal flip, radius, tile rect, SH colour)
      auxiliary.h:67-77   getRect                 (truncating tile rectangle)
      rasterizer_impl.cu:70-111, :305-310         (per-tile list ordered by (float32 view depth, index))
      forward.cu:256-445  renderCUDA              (per-pixel front-to-back compositing, all thresholds)
and relies on autograd for the backward pass, with the reference's deliberate deviations from the analytic
gradient reproduced explicitly:
  * alpha cap 0.99 is straight-through                         backward.cu:311-313, :391, :437
  * quaternion normalisation is not back-propagated            auxiliary.h:238-282
  * means2D.grad is replaced by the densification signal       backward.cu:626-629
Pinning status: the reference ships no golden vectors for this path.  The oracle is pinned
against outputs of the reference CUDA extension itself, generated on a B200 by `oracle/make_golden.py`
(fixtures under tests/golden/); see DESIGN.md §oracle.

Works in float64 (default, arbiter) or float32 (CPU baseline timing).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

TILE = 16
NEAR_N, FAR_N = 0.2, 100.0
FILTER_SIZE = 0.707106
FILTER_INV_SQUARE = 2.0
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def eval_sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """(..., 3) unit directions -> (..., (deg+1)^2) basis values (forward.cu:20-71)."""
    x, y, z = d.unbind(-1)
    b = [torch.full_like(x, SH_C0)]
    if deg > 0:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
            if deg > 2:
                b += [SH_C3[0] * y * (3.0 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4.0 * zz - xx - yy),
                      SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy), SH_C3[4] * x * (4.0 * zz - xx - yy),
                      SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3.0 * yy)]
    return torch.stack(b, dim=-1)


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """(P,4) wxyz -> (P,3,3).  Normalised in-kernel (auxiliary.h:213-235); the normalisation carries no gradient
    and the gradient w.r.t. the normalised components is handed back unchanged (auxiliary.h:238-282)."""
    s = 1.0 / q.norm(dim=-1, keepdim=True)
    qn = q + (q * s - q).detach()
    w, x, y, z = qn.unbind(-1)
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1),
    ], dim=-2)
    return R


@dataclass
class RasterOut:
    color: torch.Tensor      # (C,H,W)
    allmap: torch.Tensor     # (7,H,W)
    weight: torch.Tensor     # (P,1)
    radii: torch.Tensor      # (P,) int32
    T_render: torch.Tensor | None   # (P,3,3) node whose .grad is the render-only dL/dT (densification signal)
    depth_Tw: torch.Tensor | None   # (P,) forward T[2][2] used by the densification signal
    num_rendered: int


def preprocess(means3D, scales, rotations, viewmatrix, projmatrix, W, H, scale_modifier=1.0, transMat_precomp=None):
    """Per-surfel stage (forward.cu:148-251).  Returns dict of per-surfel tensors and the valid mask."""
    dt = means3D.dtype
    P = means3D.shape[0]
    vm = viewmatrix.to(dt)   # (4,4) = W2C^T
    pm = projmatrix.to(dt)   # (4,4) = (Proj W2C)^T
    w2c = vm.t()
    PV = pm.t()
    p_view = means3D @ w2c[:3, :3].t() + w2c[:3, 3]
    # sort keys use the float32 view depth the kernels compute (rasterizer_impl.cu:104)
    vm32, m32 = viewmatrix.float(), means3D.detach().float()
    depth32 = (m32[:, 0] * vm32[0, 2] + m32[:, 1] * vm32[1, 2] + m32[:, 2] * vm32[2, 2] + vm32[3, 2])
    in_front = p_view[:, 2] > 0.2

    if transMat_precomp is None:
        R = quat_to_rotmat(rotations)
        L0 = R[:, :, 0] * (scale_modifier * scales[:, 0:1])
        L1 = R[:, :, 1] * (scale_modifier * scales[:, 1:2])
        ones = torch.ones(P, 1, dtype=dt, device=means3D.device)
        zeros = torch.zeros(P, 1, dtype=dt, device=means3D.device)
        S = torch.stack([torch.cat([L0, zeros], -1), torch.cat([L1, zeros], -1), torch.cat([means3D, ones], -1)], dim=-1)  # (P,4,3)
        N = torch.tensor([[W / 2.0, 0, 0, (W - 1) / 2.0], [0, H / 2.0, 0, (H - 1) / 2.0], [0, 0, 0, 1.0]], dtype=dt,
                         device=means3D.device)
        T = (N @ PV) @ S                                    # (P,3,3), rows Tu, Tv, Tw
        normal = R[:, :, 2] @ w2c[:3, :3].t()
    else:
        T = transMat_precomp.reshape(P, 3, 3)
        normal = torch.tensor([0.0, 0.0, 1.0], dtype=dt, device=means3D.device).expand(P, 3)

    cosv = -(p_view * normal).sum(-1)
    mult = torch.where(cosv > 0, 1.0, -1.0).to(dt)
    normal = normal * mult[:, None].detach()

    Tu, Tv, Tw = T[:, 0], T[:, 1], T[:, 2]
    t = torch.tensor([9.0, 9.0, -1.0], dtype=dt, device=means3D.device)
    d = (t * Tw * Tw).sum(-1)
    ok = d != 0
    dsafe = torch.where(ok, d, torch.ones_like(d))
    f = t[None] / dsafe[:, None]
    center = torch.stack([(f * Tu * Tw).sum(-1), (f * Tv * Tw).sum(-1)], -1)
    h0 = center * center - torch.stack([(f * Tu * Tu).sum(-1), (f * Tv * Tv).sum(-1)], -1)
    extent = torch.sqrt(torch.clamp_min(h0, 1e-4))
    radius = torch.ceil(torch.clamp_min(extent.max(dim=-1).values, 3.0 * FILTER_SIZE)).detach()
    radius_i = torch.nan_to_num(radius, nan=0.0, posinf=2 ** 30, neginf=0).to(torch.int64)

    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    cdet = torch.nan_to_num(center.detach(), nan=0.0, posinf=1e9, neginf=-1e9)

    def trunc_div(a):  # C (int) cast truncates toward zero
        return torch.trunc(a / TILE).to(torch.int64)

    rf = radius_i.to(dt)
    rmin_x = trunc_div(cdet[:, 0] - rf).clamp(0, gx)
    rmin_y = trunc_div(cdet[:, 1] - rf).clamp(0, gy)
    rmax_x = trunc_div(cdet[:, 0] + rf + TILE - 1).clamp(0, gx)
    rmax_y = trunc_div(cdet[:, 1] + rf + TILE - 1).clamp(0, gy)
    area = (rmax_x - rmin_x) * (rmax_y - rmin_y)
    valid = in_front & (cosv != 0) & ok & (area > 0)
    radii = torch.where(valid, radius_i, torch.zeros_like(radius_i)).to(torch.int32)
    return dict(T=T, normal=normal, center=center, depth32=depth32, p_view=p_view, valid=valid, radii=radii,
                rect=(rmin_x, rmin_y, rmax_x, rmax_y))


def rasterize(means3D, opacities, scales=None, rotations=None, colors_precomp=None, shs=None, sh_degree=3, *, viewmatrix,
              projmatrix, campos, bg, W, H, scale_modifier=1.0, transMat_precomp=None, tiles=None, need_T_grad=False) -> RasterOut:
    """Forward pass of the rasterizer (differentiable through autograd).  `tiles` optionally restricts rendering to a
    list of tile indices (used by the bounded CPU baseline); pixels of other tiles stay zero."""
    dt = means3D.dtype
    dev = means3D.device
    P = means3D.shape[0]
    pre = preprocess(means3D, scales, rotations, viewmatrix, projmatrix, W, H, scale_modifier, transMat_precomp)
    T = pre["T"]
    T_r = None
    if need_T_grad:
        T_r = T * 1.0          # the node the render reads; its grad is the render-only dL/dT (backward.cu:626-629)
        T_r.retain_grad()
        T_use = T_r
    else:
        T_use = T
    valid = pre["valid"]
    if colors_precomp is None:
        dirs = means3D - campos.to(dt)[None]
        dirs = dirs / dirs.norm(dim=-1, keepdim=True)
        nb = (sh_degree + 1) ** 2
        basis = eval_sh_basis(sh_degree, dirs)                       # (P, nb)
        feat = torch.clamp_min((basis[:, :, None] * shs[:, :nb, :]).sum(1) + 0.5, 0.0)
    else:
        feat = colors_precomp
    C = feat.shape[1]
    bgv = torch.zeros(C, dtype=dt, device=dev)
    bgv[:min(C, bg.numel())] = bg.reshape(-1)[:C].to(dt)

    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rmin_x, rmin_y, rmax_x, rmax_y = pre["rect"]
    color = torch.zeros(C, H, W, dtype=dt, device=dev)
    allmap = torch.zeros(7, H, W, dtype=dt, device=dev)
    weight = torch.zeros(P, dtype=dt, device=dev)
    num_rendered = int(((rmax_x - rmin_x) * (rmax_y - rmin_y))[valid].sum())
    color_parts, allmap_parts = [], []
    tile_list = range(gx * gy) if tiles is None else tiles
    order_key = pre["depth32"].double() * 1.0   # float32 values are exactly representable
    for tile in tile_list:
        ty, tx = divmod(tile, gx)
        memb = valid & (rmin_x <= tx) & (tx < rmax_x) & (rmin_y <= ty) & (ty < rmax_y)
        ids = torch.nonzero(memb, as_tuple=False)[:, 0]
        x0, y0 = tx * TILE, ty * TILE
        x1, y1 = min(x0 + TILE, W), min(y0 + TILE, H)
        pys, pxs = torch.meshgrid(torch.arange(y0, y1, device=dev), torch.arange(x0, x1, device=dev), indexing="ij")
        pxf, pyf = pxs.reshape(-1).to(dt), pys.reshape(-1).to(dt)
        npix = pxf.shape[0]
        if ids.numel() == 0:
            col = bgv[:, None].expand(C, npix)
            color[:, y0:y1, x0:x1] = col.reshape(C, y1 - y0, x1 - x0)
            continue
        # stable order by (depth bits, index): positive floats order like their bit patterns
        ord_ = torch.argsort(order_key[ids], stable=True)
        ids = ids[ord_]
        Tt = T_use[ids]
        Tu, Tv, Tw = Tt[:, 0], Tt[:, 1], Tt[:, 2]
        k = pxf[:, None, None] * Tw[None] - Tu[None]                    # (npix, L, 3)
        l = pyf[:, None, None] * Tw[None] - Tv[None]
        p = torch.linalg.cross(k, l, dim=-1)
        pz_ok = p[..., 2] != 0
        pz = torch.where(pz_ok, p[..., 2], torch.ones_like(p[..., 2]))
        sx, sy = p[..., 0] / pz, p[..., 1] / pz
        rho3d = sx * sx + sy * sy
        cxy = pre["center"][ids]
        dx, dy = cxy[None, :, 0] - pxf[:, None], cxy[None, :, 1] - pyf[:, None]
        rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy)
        use3d = rho3d <= rho2d
        rho = torch.minimum(rho3d, rho2d)
        depth = torch.where(use3d, (sx * Tw[None, :, 0] + sy * Tw[None, :, 1]) + Tw[None, :, 2], Tw[None, :, 2].expand_as(rho))
        power = -0.5 * rho
        opa = opacities[ids, 0][None]
        oG = opa * torch.exp(power)
        alpha = oG + (torch.clamp(oG, max=0.99) - oG).detach()         # straight-through cap
        cand = pz_ok & (depth >= NEAR_N) & (power <= 0) & (alpha >= 1.0 / 255.0)
        one_m = torch.where(cand, 1.0 - alpha, torch.ones_like(alpha))
        Tincl = torch.cumprod(one_m, dim=1)
        Tbefore = torch.cat([torch.ones(npix, 1, dtype=dt, device=dev), Tincl[:, :-1]], dim=1)
        stop = cand & (Tincl.detach() < 0.0001)
        alive = torch.cumsum(stop.to(torch.int32), dim=1) == 0           # entries strictly before the first stop
        contrib = cand & alive
        w = torch.where(contrib, alpha * Tbefore, torch.zeros_like(alpha))
        Tfinal = torch.where(contrib, one_m, torch.ones_like(one_m)).prod(dim=1)

        dsafe = torch.where(contrib, depth, torch.ones_like(depth))
        m = FAR_N / (FAR_N - NEAR_N) * (1.0 - NEAR_N / dsafe)
        mw, mmw = m * w, m * m * w
        M1_ex = torch.cumsum(mw, dim=1) - mw
        M2_ex = torch.cumsum(mmw, dim=1) - mmw
        A_ex = 1.0 - Tbefore
        distortion = ((m * m * A_ex + M2_ex - 2.0 * m * M1_ex) * w).sum(1)
        Dacc = (dsafe * w).sum(1)
        Nacc = w @ pre["normal"][ids]                                    # (npix,3)
        Cacc = w @ feat[ids]                                             # (npix,C)
        med_mask = contrib & (Tbefore.detach() > 0.5)
        L = ids.shape[0]
        pos = torch.arange(1, L + 1, device=dev)[None].expand(npix, L)
        last_med = torch.where(med_mask, pos, torch.zeros_like(pos)).max(dim=1).values      # 0 => none
        med_depth = torch.where(last_med > 0, torch.gather(dsafe, 1, (last_med - 1).clamp_min(0)[:, None])[:, 0],
                                torch.zeros(npix, dtype=dt, device=dev))
        weight = weight.index_add(0, ids, w.sum(0).detach())

        col = (Cacc + Tfinal[:, None] * bgv[None]).t()
        hh, ww = y1 - y0, x1 - x0
        color[:, y0:y1, x0:x1] = col.reshape(C, hh, ww)
        allmap[0, y0:y1, x0:x1] = Dacc.reshape(hh, ww)
        allmap[1, y0:y1, x0:x1] = (1.0 - Tfinal).reshape(hh, ww)
        allmap[2:5, y0:y1, x0:x1] = Nacc.t().reshape(3, hh, ww)
        allmap[5, y0:y1, x0:x1] = med_depth.reshape(hh, ww)
        allmap[6, y0:y1, x0:x1] = distortion.reshape(hh, ww)

    return RasterOut(color, allmap, weight[:, None], pre["radii"], T_r, T[:, 2, 2], num_rendered)


def densify_signal(out: RasterOut, tanfovx: float, tanfovy: float, W: int, H: int) -> torch.Tensor:
    """means2D.grad as the reference leaves it (backward.cu:607-608, :626-629); call after .backward()."""
    focal_x = torch.tensor(W / (2.0 * tanfovx), dtype=torch.float32)
    focal_y = torch.tensor(H / (2.0 * tanfovy), dtype=torch.float32)
    Wb = int(focal_x * torch.tensor(tanfovx, dtype=torch.float32) * 2)
    Hb = int(focal_y * torch.tensor(tanfovy, dtype=torch.float32) * 2)
    g = out.T_render.grad
    depth = out.depth_Tw.detach()
    sig = torch.zeros(g.shape[0], 3, dtype=g.dtype, device=g.device)
    vis = out.radii > 0
    sig[:, 0] = torch.where(vis, g[:, 0, 2] * depth * 0.5 * Wb, torch.zeros_like(depth))
    sig[:, 1] = torch.where(vis, g[:, 1, 2] * depth * 0.5 * Hb, torch.zeros_like(depth))
    return sig


def backward_image_size(tanfovx: float, tanfovy: float, W: int, H: int) -> tuple[int, int]:
    """The (W,H) the reference backward recomputes in float32 (backward.cu:607-608); may be off by one."""
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    focal_x = f32(W) / (f32(2.0) * f32(tanfovx))
    focal_y = f32(H) / (f32(2.0) * f32(tanfovy))
    return int(focal_x * f32(tanfovx) * 2), int(focal_y * f32(tanfovy) * 2)
