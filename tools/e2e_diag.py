import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envgs_b200 import scene
from envgs_b200.pipeline import EnvGSRenderer
from envgs_b200.dist import ViewParallelTrainer
import bench
Pb, Pe, H, W = bench.WORKLOADS["config3"]
dev = torch.device("cuda")
r = EnvGSRenderer(scene.make_base_set(Pb, seed=0, s0=0.005), scene.make_env_set(Pe, seed=1, s0=0.005), device=dev)
tr = ViewParallelTrainer(r, lr=1e-6)
cams_cpu = scene.ring_cameras(16, H, W, seed=2)
cams = [c.to(dev) for c in cams_cpu]
target_host = torch.rand(H, W, 3).pin_memory(); target = target_host.to(dev)
for s in range(5): tr.step(cams[s % 16], target)
torch.cuda.synchronize()
def run(name, fn, n=15):
    torch.cuda.synchronize(); t = time.perf_counter()
    for s in range(n): fn(s)
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter() - t) / n * 1e3:.2f} ms/step")
run("resident", lambda s: tr.step(cams[s % 16], target))
run("resident + loss.item()", lambda s: tr.step(cams[s % 16], target).item())
run("resident + sync", lambda s: (tr.step(cams[s % 16], target), torch.cuda.synchronize()))
run("h2d target", lambda s: tr.step(cams[s % 16], target_host.to(dev, non_blocking=True)))
run("h2d target + loss to cpu", lambda s: tr.step(cams[s % 16], target_host.to(dev, non_blocking=True)).to("cpu"))
run("cam.to(dev) each step", lambda s: tr.step(cams_cpu[s % 16].to(dev), target))
def cam_host_pack(cam):
    return torch.cat([cam.viewmatrix.reshape(-1), cam.projmatrix.reshape(-1), cam.campos.reshape(-1), cam.K.reshape(-1), cam.R.reshape(-1), cam.T.reshape(-1)]).pin_memory()
def cam_from_pack(cam, pd):
    return scene.Camera(cam.H, cam.W, pd[35:44].view(3, 3), pd[44:53].view(3, 3), pd[53:56].view(3, 1), cam.tanfovx, cam.tanfovy, pd[0:16].view(4, 4), pd[16:32].view(4, 4), pd[32:35], cam.znear, cam.zfar)
packs = [cam_host_pack(c) for c in cams_cpu]
run("pack cam", lambda s: tr.step(cam_from_pack(cams_cpu[s % 16], packs[s % 16].to(dev, non_blocking=True)), target))
run("pack cam + h2d target + loss to cpu", lambda s: tr.step(cam_from_pack(cams_cpu[s % 16], packs[s % 16].to(dev, non_blocking=True)), target_host.to(dev, non_blocking=True)).to("cpu"))
run("resident + loss.item() again", lambda s: tr.step(cams[s % 16], target).item())
run("resident again", lambda s: tr.step(cams[s % 16], target))
