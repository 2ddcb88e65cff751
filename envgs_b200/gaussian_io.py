"""On-disk formats and the camera feed either side of the hot path.

* `save_ply` / `load_ply` — the reference's per-set point-cloud file, byte-compatible with what
  `GaussianModel.save_ply` writes through `plyfile` (easyvolcap/utils/gaussian2d_utils.py:918-961): binary little-endian
  PLY, one `vertex` element, float32 properties in the order x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*, SH
  coefficients channel-major (`transpose(1, 2).flatten(1)`), raw (pre-activation) opacity / scale values.  `load_ply`
  follows :963-1000 (property lookup by name, f_rest / scale / rot sorted by numeric suffix).  `plyfile` itself is not
  needed: the writer emits exactly its header text and row packing.
* `save_checkpoint` / `load_checkpoint` — the `.pt` the reference's trainer writes (easyvolcap/utils/net_utils.py:481-523):
  `{'model': state_dict, 'epoch': int[, 'optimizer': ...]}` with the EnvGS sampler's key names (`sampler.pcd._xyz`, ...,
  `sampler.env._rotation`, buffers `active_sh_degree`, `max_radii2D`, `xyz_gradient_accum`, `denom`, `xyz_weight_accum`;
  envgs_sampler.py:165, gaussian2d_utils.py:294-312, :440-467).
* `camera_from_batch` — the per-view camera block the two operators are fed with, from EasyVolcap's batch fields
  (K, R, T, H, W, n, f): `prepare_gaussian_camera` (gaussian2d_utils.py:67-100) with `getProjectionMatrix` (:40-64).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import scene

PCD_PREFIX = "sampler.pcd."
ENV_PREFIX = "sampler.env."
PARAM_NAMES = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_specular", "_roughness")


@dataclass
class RawSurfels:
    """One Gaussian set in the reference's parameter space (pre-activation), names as in its state dict."""
    _xyz: torch.Tensor             # (P,3)
    _features_dc: torch.Tensor     # (P,1,3)
    _features_rest: torch.Tensor   # (P,(D+1)^2-1,3)
    _scaling: torch.Tensor         # (P,2)  log
    _rotation: torch.Tensor        # (P,4)  wxyz, unnormalised
    _opacity: torch.Tensor         # (P,1)  logit
    _specular: torch.Tensor | None = None    # (P,1) logit
    _roughness: torch.Tensor | None = None   # (P,1) logit
    active_sh_degree: int = 3
    buffers: dict = field(default_factory=dict)   # max_radii2D, xyz_gradient_accum, denom, xyz_weight_accum

    @property
    def P(self) -> int:
        return self._xyz.shape[0]

    @property
    def max_sh_degree(self) -> int:
        return int(round(math.sqrt(self._features_rest.shape[1] + 1))) - 1

    @staticmethod
    def from_activated(s: scene.SurfelSet, eps: float = 1e-6) -> "RawSurfels":
        lg = lambda t: torch.logit(t.clamp(eps, 1 - eps))
        return RawSurfels(s.means.clone(), s.shs[:, :1].clone(), s.shs[:, 1:].clone(), torch.log(s.scales), s.rotations.clone(),
                          lg(s.opacities), lg(s.specular), lg(s.roughness))

    def activate(self) -> scene.SurfelSet:
        """What `render()` / `HardwareRendering` hand to the operators (gaussian2d_utils.py:362-392): exp scales, sigmoid
        opacity / specular / roughness, normalised rotations, dc + rest SH concatenated."""
        P = self.P
        dev = self._xyz.device
        spec = torch.sigmoid(self._specular) if self._specular is not None else torch.full((P, 1), 1e-3, device=dev)
        rough = torch.sigmoid(self._roughness) if self._roughness is not None else torch.full((P, 1), 0.5, device=dev)
        return scene.SurfelSet(means=self._xyz, scales=torch.exp(self._scaling),
                               rotations=torch.nn.functional.normalize(self._rotation, dim=-1),
                               opacities=torch.sigmoid(self._opacity),
                               shs=torch.cat([self._features_dc, self._features_rest], dim=1).contiguous(),
                               specular=spec, roughness=rough)


def ply_attribute_names(n_dc: int, n_rest: int, n_scale: int, n_rot: int) -> list[str]:
    """construct_list_of_attributes (gaussian2d_utils.py:918-933)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(raw: RawSurfels, path: str, bounds: torch.Tensor | None = None) -> int:
    """Write the set (optionally only the points inside `bounds` = (min xyz, max xyz)); returns the vertex count."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    xyz = raw._xyz.detach()
    mask = ((xyz >= bounds[0]) & (xyz <= bounds[1])).all(dim=-1) if bounds is not None else torch.ones(xyz.shape[0], dtype=torch.bool, device=xyz.device)
    c = lambda t: t.detach()[mask].cpu().float()
    xyz_ = c(raw._xyz)
    cols = [xyz_, torch.zeros_like(xyz_),
            c(raw._features_dc).transpose(1, 2).flatten(start_dim=1), c(raw._features_rest).transpose(1, 2).flatten(start_dim=1),
            c(raw._opacity), c(raw._scaling), c(raw._rotation)]
    table = torch.cat(cols, dim=1).contiguous().numpy().astype("<f4")
    names = ply_attribute_names(cols[2].shape[1], cols[3].shape[1], cols[5].shape[1], cols[6].shape[1])
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % table.shape[0]
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(table.tobytes(order="C"))
    return table.shape[0]


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
              "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_ply_vertices(path: str) -> dict:
    """Minimal PLY reader for the first element (scalar properties only): binary little/big endian or ascii."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_first = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is None:
                    count, in_first = int(tok[2]), True
                else:
                    in_first = False
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=count, ndmin=2)
            return {n: data[:, i] for i, (n, _) in enumerate(props)}
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, order + t) for n, t in props])
        arr = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
        return {n: arr[n] for n, _ in props}


def load_ply(path: str, max_sh_degree: int = 3, device: str | torch.device = "cpu") -> RawSurfels:
    v = read_ply_vertices(path)
    num = lambda prefix: sorted((n for n in v if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1)
    P = xyz.shape[0]
    f_dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], axis=1).reshape(P, 3, 1)
    rest_names = num("f_rest_")
    if len(rest_names) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{path}: {len(rest_names)} f_rest_* properties, expected {3 * (max_sh_degree + 1) ** 2 - 3}")
    f_rest = np.stack([v[n] for n in rest_names], axis=1).reshape(P, 3, (max_sh_degree + 1) ** 2 - 1) if rest_names else np.zeros((P, 3, 0))
    scales = np.stack([v[n] for n in num("scale_")], axis=1)
    rots = np.stack([v[n] for n in num("rot")], axis=1)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=device)
    return RawSurfels(t(xyz), t(f_dc).transpose(1, 2).contiguous(), t(f_rest).transpose(1, 2).contiguous(), t(scales), t(rots),
                      t(v["opacity"])[:, None], active_sh_degree=max_sh_degree)


# ---- checkpoint ---------------------------------------------------------------------------------------------------
def state_dict_of(sets: dict) -> dict:
    """{prefix: RawSurfels} -> flat state dict with the reference's key names."""
    sd = {}
    for prefix, raw in sets.items():
        for name in PARAM_NAMES:
            t = getattr(raw, name)
            if t is not None:
                sd[prefix + name] = t.detach().clone()
        sd[prefix + "active_sh_degree"] = torch.full((1,), int(raw.active_sh_degree), dtype=torch.long)
        P = raw.P
        defaults = dict(max_radii2D=torch.zeros(P), xyz_gradient_accum=torch.zeros(P, 1), denom=torch.zeros(P, 1),
                        xyz_weight_accum=torch.zeros(P, 1))
        for k, dflt in defaults.items():
            sd[prefix + k] = raw.buffers.get(k, dflt).detach().clone()
    return sd


def save_checkpoint(model_dir: str, sets: dict, epoch: int, optimizer_state: dict | None = None, latest: bool = False,
                    save_lim: int = 5) -> str:
    """net_utils.save_model (:481-523): `<epoch>.pt` (or latest.pt), keeping at most `save_lim` numbered files."""
    os.makedirs(model_dir, exist_ok=True)
    blob = {"model": state_dict_of(sets), "epoch": int(epoch)}
    if optimizer_state is not None:
        blob["optimizer"] = optimizer_state
    path = os.path.join(model_dir, "latest.pt" if latest else f"{epoch}.pt")
    torch.save(blob, path)
    pts = [int(p.split(".")[0]) for p in os.listdir(model_dir) if p != "latest.pt" and p.endswith(".pt") and p.split(".")[0].isnumeric()]
    if len(pts) > save_lim:
        os.remove(os.path.join(model_dir, f"{min(pts)}.pt"))
    return path


def load_checkpoint(path: str, device: str | torch.device = "cpu", prefixes=(PCD_PREFIX, ENV_PREFIX)):
    """-> ({prefix: RawSurfels}, epoch, optimizer_state or None).  Point counts come from the file (the reference resizes its
    parameters in a load_state_dict pre-hook, gaussian2d_utils.py:469-476)."""
    blob = torch.load(path, map_location=device)
    sd = blob["model"]
    sets = {}
    for prefix in prefixes:
        if prefix + "_xyz" not in sd:
            continue
        kw = {n: sd.get(prefix + n) for n in PARAM_NAMES}
        raw = RawSurfels(**kw)
        if prefix + "active_sh_degree" in sd:
            raw.active_sh_degree = int(sd[prefix + "active_sh_degree"].item())
        raw.buffers = {k: sd[prefix + k] for k in ("max_radii2D", "xyz_gradient_accum", "denom", "xyz_weight_accum") if prefix + k in sd}
        sets[prefix] = raw
    return sets, int(blob.get("epoch", -1)), blob.get("optimizer")


# ---- camera feed --------------------------------------------------------------------------------------------------
def camera_from_batch(batch) -> scene.Camera:
    """EasyVolcap batch (fields K (B,3,3), R (B,3,3), T (B,3,1), H, W, n, f; attribute or key access; first view is used, like
    the reference) -> the camera block of both operators.  The principal point is ignored exactly as in the reference
    (FoV from the focal length only, gaussian2d_utils.py:79-82)."""
    get = (lambda k: batch[k]) if isinstance(batch, dict) else (lambda k: getattr(batch, k))
    first = lambda v: v[0] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] >= 1 and v.dim() in (1, 3) else v
    K, R, T = (first(get(k)).float() for k in ("K", "R", "T"))
    scalar = lambda v: float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)
    H, W, n, f = int(scalar(get("H"))), int(scalar(get("W"))), scalar(get("n")), scalar(get("f"))
    fov = lambda focal, pixels: 2 * math.atan(pixels / (2 * focal))       # focal2fov
    fovx, fovy = fov(float(K[0, 0]), W), fov(float(K[1, 1]), H)
    tanfovx, tanfovy = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    dev = R.device
    w2c = torch.eye(4, device=dev)
    w2c[:3, :3] = R
    w2c[:3, 3:] = T.reshape(3, 1)
    proj = scene.get_projection_matrix(tanfovx, tanfovy, n, f).to(dev)
    view = w2c.t().contiguous()
    full = (view @ proj.t()).contiguous()
    campos = (-R.t() @ T.reshape(3, 1))[:, 0].contiguous()
    return scene.Camera(H=H, W=W, K=K, R=R, T=T.reshape(3, 1), tanfovx=tanfovx, tanfovy=tanfovy, viewmatrix=view, projmatrix=full,
                        campos=campos, znear=n, zfar=f)
