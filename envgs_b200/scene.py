"""Deterministic synthetic scenes and cameras for parity tests and benchmarks.

Everything is generated on the CPU with seeded `torch.Generator`s in fp32 and moved to the device by the caller, so
tests, the oracle, the reference build and the benchmark all see bit-identical inputs.  Camera conventions follow
EasyVolcap's `prepare_gaussian_camera` (easyvolcap/utils/gaussian2d_utils.py:40-100): `viewmatrix = W2C^T`,
`projmatrix = (Proj @ W2C)^T`, principal point ignored.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class SurfelSet:
    means: torch.Tensor       # (P,3)
    scales: torch.Tensor      # (P,2)  already exp-activated
    rotations: torch.Tensor   # (P,4)  wxyz, unit
    opacities: torch.Tensor   # (P,1)  already sigmoid-activated
    shs: torch.Tensor         # (P,16,3)
    specular: torch.Tensor    # (P,1)
    roughness: torch.Tensor   # (P,1)

    def to(self, device):
        return SurfelSet(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])

    @property
    def P(self):
        return self.means.shape[0]


@dataclass
class Camera:
    H: int
    W: int
    K: torch.Tensor            # (3,3)
    R: torch.Tensor            # (3,3) world -> camera
    T: torch.Tensor            # (3,1)
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # (4,4) W2C^T
    projmatrix: torch.Tensor   # (4,4) (Proj @ W2C)^T
    campos: torch.Tensor       # (3,)
    znear: float = 0.2
    zfar: float = 100.0

    def to(self, device):
        kw = {}
        for f in self.__dataclass_fields__:
            v = getattr(self, f)
            kw[f] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Camera(**kw)


def _sh_colors(g, P):
    dc = (torch.rand(P, 1, 3, generator=g) - 0.5) / 0.28209479177387814
    rest = 0.05 * torch.randn(P, 15, 3, generator=g)
    return torch.cat([dc, rest], dim=1).contiguous()


def make_base_set(P: int, seed: int = 0, s0: float = 0.005) -> SurfelSet:
    """Half volume clutter in [-2,2]^3, half on a unit sphere + 6 random planes (so depth/normal maps and hence
    reflection rays are non-degenerate)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_vol = P // 2
    n_surf = P - n_vol
    vol = (torch.rand(n_vol, 3, generator=g) * 4.0 - 2.0)
    n_sph = n_surf // 2
    sph = torch.randn(n_sph, 3, generator=g)
    sph = sph / sph.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    n_pl = n_surf - n_sph
    planes_n = torch.randn(6, 3, generator=g)
    planes_n = planes_n / planes_n.norm(dim=-1, keepdim=True)
    planes_d = torch.rand(6, generator=g) * 1.5 + 0.25
    which = torch.randint(0, 6, (n_pl,), generator=g)
    q = torch.rand(n_pl, 3, generator=g) * 4.0 - 2.0
    nrm = planes_n[which]
    q = q - ((q * nrm).sum(-1, keepdim=True) - planes_d[which][:, None]) * nrm   # project onto the plane
    q = q.clamp(-2.0, 2.0)
    surf = torch.cat([sph, q], dim=0) + 0.01 * torch.randn(n_surf, 3, generator=g)
    means = torch.cat([vol, surf], dim=0)
    perm = torch.randperm(P, generator=g)
    means = means[perm].contiguous()
    scales = torch.exp(math.log(s0) + 0.5 * torch.randn(P, 2, generator=g))
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    shs = _sh_colors(g, P)
    spec = torch.sigmoid(torch.randn(P, 1, generator=g) - 2.0)
    rough = torch.full((P, 1), 0.5)
    return SurfelSet(means.float(), scales.float(), rot.float(), opac.float(), shs.float(), spec.float(), rough.float())


def make_env_set(P: int, seed: int = 1, s0: float = 0.005) -> SurfelSet:
    """Environment surfels on shells r in [8, 20] (inside env_bounds +-20), scales x8."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = torch.randn(P, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    r = 8.0 + 12.0 * torch.rand(P, 1, generator=g)
    means = (d * r).contiguous()
    scales = torch.exp(math.log(s0 * 8.0) + 0.5 * torch.randn(P, 2, generator=g))
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    shs = _sh_colors(g, P)
    spec = torch.sigmoid(torch.randn(P, 1, generator=g) - 2.0)
    rough = torch.full((P, 1), 0.5)
    return SurfelSet(means.float(), scales.float(), rot.float(), opac.float(), shs.float(), spec.float(), rough.float())


def get_projection_matrix(tanfovx: float, tanfovy: float, znear: float, zfar: float) -> torch.Tensor:
    P = torch.zeros(4, 4)
    P[0, 0] = 1.0 / tanfovx
    P[1, 1] = 1.0 / tanfovy
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(H: int, W: int, R: torch.Tensor, T: torch.Tensor, fx: float, fy: float, znear: float = 0.2,
                zfar: float = 100.0) -> Camera:
    K = torch.tensor([[fx, 0.0, W / 2.0], [0.0, fy, H / 2.0], [0.0, 0.0, 1.0]])
    fovx = 2.0 * math.atan(W / (2.0 * fx))
    fovy = 2.0 * math.atan(H / (2.0 * fy))
    tanfovx, tanfovy = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    w2c = torch.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3:] = T
    view = w2c.t().contiguous()
    proj = get_projection_matrix(tanfovx, tanfovy, znear, zfar).t()
    full = (view @ proj).contiguous()
    campos = (-R.t() @ T)[:, 0].contiguous()
    return Camera(H, W, K.float(), R.float().contiguous(), T.float().contiguous(), tanfovx, tanfovy, view.float(), full.float(),
                  campos.float(), znear, zfar)


def look_at_camera(H: int, W: int, eye: torch.Tensor, target: torch.Tensor | None = None, focal_scale: float = 0.78) -> Camera:
    target = torch.zeros(3) if target is None else target
    fwd = target - eye
    fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, 0.0, 1.0])
    if abs(float(fwd @ up)) > 0.99:
        up = torch.tensor([0.0, 1.0, 0.0])
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], dim=0)   # rows: camera x (right), y (down), z (forward)
    T = -(R @ eye)[:, None]
    f = focal_scale * W
    return make_camera(H, W, R, T, f, f)


def ring_cameras(n: int, H: int, W: int, seed: int = 2, radius: float = 4.0) -> list[Camera]:
    """Look-at-origin cameras on a ring r = radius +- 0.5, elevation U(-20deg, 40deg), fx = fy = 0.78 W."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cams = []
    for i in range(n):
        az = 2.0 * math.pi * (i + float(torch.rand(1, generator=g))) / n
        el = math.radians(-20.0 + 60.0 * float(torch.rand(1, generator=g)))
        r = radius - 0.5 + float(torch.rand(1, generator=g))
        eye = torch.tensor([r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el)])
        cams.append(look_at_camera(H, W, eye))
    return cams


def camera_rays(cam: Camera) -> tuple[torch.Tensor, torch.Tensor]:
    """(H,W,3) origins / un-normalised directions with camera-space z = 1 and half-pixel centres
    (easyvolcap/utils/ray_utils.py:24-78, z_depth=True, correct_pix=True)."""
    dev = cam.K.device
    i, j = torch.meshgrid(torch.arange(cam.H, dtype=torch.float32, device=dev),
                          torch.arange(cam.W, dtype=torch.float32, device=dev), indexing="ij")
    xy1 = torch.stack([j + 0.5, i + 0.5, torch.ones_like(i)], dim=-1)
    pix_cam = xy1 @ torch.linalg.inv(cam.K).t()
    pix_world = (pix_cam - cam.T[:, 0]) @ cam.R          # R^T (x - T) written row-wise
    ray_o = cam.campos.expand_as(pix_world)
    ray_d = pix_world - ray_o
    return ray_o.contiguous(), ray_d.contiguous()


def quad_vertices(means: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Surfel -> 4 quad corners at +-3 sigma in the reference order [(-,+),(-,-),(+,+),(+,-)] and the two triangles
    per quad (easyvolcap/utils/optix_utils.py:39-69)."""
    q = rotations / rotations.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    tu = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], dim=-1)
    tv = torch.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], dim=-1)
    su, sv = scales[:, :1], scales[:, 1:2]
    sig = torch.tensor([[-1.0, 1.0], [-1.0, -1.0], [1.0, 1.0], [1.0, -1.0]], device=means.device) * 3.0
    v = means[:, None, :] + sig[None, :, :1] * (su * tu)[:, None, :] + sig[None, :, 1:2] * (sv * tv)[:, None, :]
    v = v.reshape(-1, 3).contiguous()
    idx = torch.arange(v.shape[0], device=means.device).reshape(-1, 4)
    f = torch.stack([idx[:, :3], idx[:, 1:]], dim=1).reshape(-1, 3).int().contiguous()
    return v, f
