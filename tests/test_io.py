"""CPU: on-disk formats and the camera feed either side of the hot path."""
import struct
import types

import numpy as np
import torch

from envgs_b200 import gaussian_io as IO
from envgs_b200 import scene


def _raw(P=37, seed=0):
    return IO.RawSurfels.from_activated(scene.make_base_set(P, seed=seed, s0=0.04))


def test_ply_bytes_are_what_plyfile_writes(tmp_path):
    """Header text and row packing restated independently of the writer: plyfile's `PlyData([PlyElement.describe(arr,
    'vertex')]).write()` of an all-'f4' structured array (gaussian2d_utils.py:952-961)."""
    raw = _raw()
    path = str(tmp_path / "sub" / "pcd.ply")
    assert IO.save_ply(raw, path) == raw.P
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + \
        ["opacity", "scale_0", "scale_1", "rot_0", "rot_1", "rot_2", "rot_3"]
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % raw.P + "".join(f"property float {n}\n" for n in names) + "end_header\n"
    rows = b""
    for i in range(raw.P):
        vals = list(raw._xyz[i]) + [0.0, 0.0, 0.0] + list(raw._features_dc[i].t().reshape(-1)) + list(raw._features_rest[i].t().reshape(-1)) + \
            list(raw._opacity[i]) + list(raw._scaling[i]) + list(raw._rotation[i])
        rows += struct.pack("<%df" % len(names), *[float(v) for v in vals])
    assert open(path, "rb").read() == header.encode("ascii") + rows


def test_ply_round_trip_and_bounds(tmp_path):
    raw = _raw(101, seed=3)
    path = str(tmp_path / "a.ply")
    IO.save_ply(raw, path)
    back = IO.load_ply(path, max_sh_degree=3)
    for n in ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity"):
        assert torch.equal(getattr(back, n), getattr(raw, n)), n
    assert back._features_dc.shape == (101, 1, 3) and back._features_rest.shape == (101, 15, 3)
    bounds = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    n_in = IO.save_ply(raw, path, bounds=bounds)
    mask = ((raw._xyz >= bounds[0]) & (raw._xyz <= bounds[1])).all(-1)
    assert n_in == int(mask.sum()) and torch.equal(IO.load_ply(path)._xyz, raw._xyz[mask])
    # activation is what the operators are fed with
    act = back.activate()
    assert torch.allclose(act.opacities, torch.sigmoid(raw._opacity)) and torch.allclose(act.scales, torch.exp(raw._scaling))
    assert act.shs.shape == (101, 16, 3)


def test_ply_reader_accepts_ascii_and_other_property_orders(tmp_path):
    path = str(tmp_path / "ascii.ply")
    names = ["x", "y", "z", "opacity", "rot_3", "rot_2", "rot_1", "rot_0", "scale_1", "scale_0", "f_dc_0", "f_dc_1", "f_dc_2"]
    rows = np.arange(2 * len(names), dtype=np.float64).reshape(2, -1)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in names) + "end_header\n")
        for r in rows:
            f.write(" ".join(str(v) for v in r) + "\n")
    raw = IO.load_ply(path, max_sh_degree=0)
    assert raw._rotation[0].tolist() == [7.0, 6.0, 5.0, 4.0] and raw._scaling[1].tolist() == [22.0, 21.0]
    assert raw._features_rest.shape == (2, 0, 3)


def test_checkpoint_round_trip_uses_reference_key_names(tmp_path):
    pcd, env = _raw(50, 1), _raw(20, 2)
    pcd.buffers["denom"] = torch.arange(50.0)[:, None]
    d = str(tmp_path / "trained_model")
    p = IO.save_checkpoint(d, {IO.PCD_PREFIX: pcd, IO.ENV_PREFIX: env}, epoch=7, optimizer_state={"step": 3})
    blob = torch.load(p)
    assert set(blob) == {"model", "epoch", "optimizer"} and blob["epoch"] == 7
    for k in ("sampler.pcd._xyz", "sampler.pcd._features_rest", "sampler.env._rotation", "sampler.env._specular",
              "sampler.pcd.active_sh_degree", "sampler.env.xyz_gradient_accum", "sampler.pcd.max_radii2D"):
        assert k in blob["model"], k
    sets, epoch, opt = IO.load_checkpoint(p)
    assert epoch == 7 and opt == {"step": 3}
    assert torch.equal(sets[IO.PCD_PREFIX]._xyz, pcd._xyz) and torch.equal(sets[IO.ENV_PREFIX]._roughness, env._roughness)
    assert torch.equal(sets[IO.PCD_PREFIX].buffers["denom"], pcd.buffers["denom"])
    for e in range(8, 15):   # save_lim = 5 numbered files are kept
        IO.save_checkpoint(d, {IO.PCD_PREFIX: pcd}, epoch=e)
    import os
    assert sorted(f for f in os.listdir(d) if f.endswith(".pt")) == ["10.pt", "11.pt", "12.pt", "13.pt", "14.pt"]


def test_camera_from_batch_matches_scene_camera():
    cam = scene.ring_cameras(3, 90, 160, seed=4)[1]
    batch = types.SimpleNamespace(K=cam.K[None], R=cam.R[None], T=cam.T[None], H=torch.tensor([90]), W=torch.tensor([160]),
                                  n=torch.tensor([0.2]), f=torch.tensor([100.0]))
    c2 = IO.camera_from_batch(batch)
    assert (c2.H, c2.W) == (90, 160)
    for k in ("viewmatrix", "projmatrix", "campos"):
        assert torch.allclose(getattr(c2, k), getattr(cam, k), atol=1e-6), k
    assert abs(c2.tanfovx - cam.tanfovx) < 1e-7 and abs(c2.tanfovy - cam.tanfovy) < 1e-7
    c3 = IO.camera_from_batch(dict(K=cam.K, R=cam.R, T=cam.T, H=90, W=160, n=0.2, f=100.0))
    assert torch.allclose(c3.projmatrix, cam.projmatrix, atol=1e-6)
