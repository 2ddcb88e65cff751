"""Development aid (torchrun, N GPUs): the overlapped / sharded exchange of ViewParallelTrainer against a single-process
emulation of the same N views (gradients summed, plain full-size Adam)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from envgs_b200 import scene
from envgs_b200.dist import ViewParallelTrainer, init_process_group_from_env, FlatAdam, FlatParameters
from envgs_b200.pipeline import EnvGSRenderer

rank, local, world = init_process_group_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
H, W = 270, 480
base, env = scene.make_base_set(20000, seed=0, s0=0.02), scene.make_env_set(8192, seed=1, s0=0.02)
cams = [c.to(dev) for c in scene.ring_cameras(2 * world, H, W, seed=2)]
target = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(3)).to(dev)
r = EnvGSRenderer(base, env, device=dev)
tr = ViewParallelTrainer(r, lr=1e-3)
for step in range(2):
    tr.step(cams[step * world + rank], target)
torch.cuda.synchronize()
mine = {n: p.detach().clone() for n, p in r.named_parameters()}
if rank == 0:
    r2 = EnvGSRenderer(base, env, device=dev)
    fp = FlatParameters(r2, pad=1024 * 8)
    adam = FlatAdam(fp.flat, fp.flat_grad, lr=1e-3)
    for step in range(2):
        fp.zero_grad()
        for k in range(world):
            out = r2(cams[step * world + k], target)
            r2.loss(out, target).backward()
        fp.collect_stray_grads()
        adam.step(grad_scale=1.0 / world, zero_grad=True)
    torch.cuda.synchronize()
    r0 = EnvGSRenderer(base, env, device=dev)
    init = dict(r0.named_parameters())
    ref = dict(r2.named_parameters())
    worst = 0.0
    for n in mine:
        d = (mine[n] - ref[n]).norm()
        upd = (ref[n] - init[n]).norm().clamp_min(1e-30)
        worst = max(worst, float(d / upd))
        print(f"  {n:16s} rel diff of the update {float(d / upd):.3e}")
    print(f"world {world}: shard={tr.shard} overlap={tr.overlap} factored={tr.factored}; worst relative difference of a parameter update {worst:.3e}")
dist.barrier()
dist.destroy_process_group()
