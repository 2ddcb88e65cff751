"""PyTorch twin of the reference's densification code, used only by tests/test_gpu_densify.py: a restatement of
easyvolcap/utils/gaussian2d_utils.py with per-tensor parameters and a real torch.optim.Adam (one param group per tensor, as
easyvolcap/runners/optimizers.py builds it).  Function by function: replace/cat/prune optimizer surgery :517-592, statistics
:661-716, clone/split :718-760, selection rules :762-836, densify_and_prune :867-899.  The only departure: `torch.normal` in
`split` is fed from a caller-supplied table of standard normals indexed by (copy, is_clone, original row) so that the device
implementation can consume the same numbers."""
import torch
import torch.nn as nn

NAMES = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_specular", "_roughness")


def build_rotation(r):
    q = r / torch.norm(r, dim=-1)[..., None]
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = torch.zeros(q.shape[:-1] + (3, 3), dtype=r.dtype, device=r.device)
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


class TwinModel:
    def __init__(self, params: dict, stats: dict, noise: torch.Tensor, spatial_scale=1.0):
        for n in NAMES:
            setattr(self, n, nn.Parameter(params[n].clone()))
        self.xyz_gradient_accum, self.denom = stats["xyz_gradient_accum"].clone()[:, None], stats["denom"].clone()[:, None]
        self.max_radii2D, self.xyz_weight_accum = stats["max_radii2D"].clone(), stats["xyz_weight_accum"].clone()[:, None]
        self.spatial_scale = spatial_scale
        self.noise = noise                                         # (n_split, 2, P0, 2)
        P = params["_xyz"].shape[0]
        dev = params["_xyz"].device
        self.src = torch.arange(P, device=dev)                     # provenance of every current row
        self.is_clone = torch.zeros(P, dtype=torch.long, device=dev)
        self.optimizer = torch.optim.Adam([{"params": [getattr(self, n)], "name": n} for n in NAMES], lr=1e-3, eps=1e-15)

    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))

    # ---- optimizer surgery (:538-592) ----
    def _prune_optimizer(self, mask):
        for g in self.optimizer.param_groups:
            st = self.optimizer.state.get(g["params"][0], None)
            if st is not None and len(st):
                st["exp_avg"], st["exp_avg_sq"] = st["exp_avg"][mask], st["exp_avg_sq"][mask]
                del self.optimizer.state[g["params"][0]]
                g["params"][0] = nn.Parameter(g["params"][0][mask].requires_grad_(True))
                self.optimizer.state[g["params"][0]] = st
            else:
                g["params"][0] = nn.Parameter(g["params"][0][mask].requires_grad_(True))
            setattr(self, g["name"], g["params"][0])

    def cat_tensors_to_optimizer(self, d):
        for g in self.optimizer.param_groups:
            ext = d[g["name"]]
            st = self.optimizer.state.get(g["params"][0], None)
            if st is not None and len(st):
                st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
                st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
                del self.optimizer.state[g["params"][0]]
                g["params"][0] = nn.Parameter(torch.cat((g["params"][0], ext), dim=0).requires_grad_(True))
                self.optimizer.state[g["params"][0]] = st
            else:
                g["params"][0] = nn.Parameter(torch.cat((g["params"][0], ext), dim=0).requires_grad_(True))
            setattr(self, g["name"], g["params"][0])

    def prune_points(self, mask):
        keep = ~mask
        self._prune_optimizer(keep)
        self.xyz_gradient_accum, self.denom = self.xyz_gradient_accum[keep], self.denom[keep]        # prune_stats :676-687
        self.max_radii2D, self.xyz_weight_accum = self.max_radii2D[keep], self.xyz_weight_accum[keep]
        self.src, self.is_clone = self.src[keep], self.is_clone[keep]

    # ---- statistics (:661-703) ----
    def get_xyz_gradient_avg(self):
        avg = self.xyz_gradient_accum / self.denom
        avg[avg.isnan()] = 0.0
        return avg

    def get_xyz_weight_avg(self):
        avg = self.xyz_weight_accum / self.denom
        avg[avg.isnan()] = 0.0
        return avg

    def densify_stats(self, sel, split=1, ratio=1.0):
        self.xyz_gradient_accum = torch.cat([self.xyz_gradient_accum, self.xyz_gradient_accum[sel].repeat(split, 1) * ratio], dim=0)
        new_w = self.xyz_weight_accum[sel].repeat(split, 1) * self.xyz_weight_accum.max()
        self.denom = torch.cat([self.denom, self.denom[sel].repeat(split, 1)], dim=0)
        self.max_radii2D = torch.cat([self.max_radii2D, self.max_radii2D[sel].repeat(split) * ratio], dim=0)
        self.xyz_weight_accum = torch.cat([self.xyz_weight_accum, new_w], dim=0)

    # ---- clone / split (:718-760) ----
    def clone(self, mask):
        self.cat_tensors_to_optimizer({n: getattr(self, n)[mask] for n in NAMES})
        self.densify_stats(mask, 1, 1.0)
        self.src = torch.cat([self.src, self.src[mask]])
        self.is_clone = torch.cat([self.is_clone, torch.ones_like(self.is_clone[mask])])

    def split(self, mask, N=2, ratio=0.8):
        stds = self.get_scaling[mask].repeat(N, 1)
        stds = torch.cat([stds, torch.zeros_like(stds[:, :1])], dim=-1)
        src, isc = self.src[mask], self.is_clone[mask]
        z = torch.cat([self.noise[k, isc, src] for k in range(N)], dim=0)                 # stands in for torch.normal(0, 1)
        samples = torch.cat([z, torch.zeros_like(z[:, :1])], dim=-1) * stds
        rots = build_rotation(self._rotation[mask]).repeat(N, 1, 1)
        new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self._xyz[mask].repeat(N, 1)
        d = {n: getattr(self, n)[mask].repeat(N, *([1] * (getattr(self, n).dim() - 1))) for n in NAMES}
        d["_xyz"] = new_xyz
        d["_scaling"] = torch.log(self.get_scaling[mask].repeat(N, 1) / (ratio * N))
        self.cat_tensors_to_optimizer(d)
        self.densify_stats(mask, N, 1.0 / (ratio * N))
        self.src = torch.cat([self.src, src.repeat(N)])
        self.is_clone = torch.cat([self.is_clone, isc.repeat(N)])
        n_split = int(mask.sum())
        self.prune_points(torch.cat((mask, torch.zeros(n_split * N, device=mask.device, dtype=torch.bool))))

    # ---- rules (:762-836, :849-899) ----
    def densify_and_prune(self, min_opacity, min_gradient, grad_thr, size_thr, split_screen_threshold=None, prune_visibility=False,
                          n_after=None):
        grads = self.get_xyz_gradient_avg()
        high = (grads >= grad_thr).squeeze(-1)
        sel = (torch.max(self.get_scaling, dim=1).values <= size_thr * self.spatial_scale) & high
        if int(sel.sum()) > 0:
            self.clone(sel)
        grads = self.get_xyz_gradient_avg()
        high = (grads >= grad_thr).squeeze(-1)
        sel = torch.max(self.get_scaling, dim=1).values > size_thr * self.spatial_scale
        if split_screen_threshold is not None:
            sel = sel | (self.max_radii2D > split_screen_threshold)
        sel = sel & high
        if int(sel.sum()) > 0:
            self.split(sel, 2)
        P = self._xyz.shape[0]
        dev = self._xyz.device
        occ = (self.get_opacity < min_opacity).squeeze(-1) if min_opacity is not None else torch.zeros(P, dtype=torch.bool, device=dev)
        if min_gradient is not None:
            grads = self.get_xyz_gradient_avg()
            mg = ((grads <= min_gradient) & (self.denom != 0)).squeeze(-1)
        else:
            mg = torch.zeros(P, dtype=torch.bool, device=dev)
        pm = occ | mg
        if int(pm.sum()) > 0:
            self.prune_points(pm)
        if prune_visibility and n_after is not None and self._xyz.shape[0] - n_after > 0:
            w = self.get_xyz_weight_avg()
            _, idx = torch.topk(w[..., 0], self._xyz.shape[0] - n_after, largest=False)
            m = torch.zeros(self._xyz.shape[0], dtype=torch.bool, device=dev)
            m[idx] = True
            self.prune_points(m)

    def moments(self, name):
        st = self.optimizer.state[getattr(self, name)]
        return st["exp_avg"], st["exp_avg_sq"]
