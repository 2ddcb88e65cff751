import torch, time
x = torch.rand(1080, 1920, 3).pin_memory()
print("pinned:", x.is_pinned())
torch.cuda.synchronize()
for n in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = x.to("cuda", non_blocking=True); b.record(); torch.cuda.synchronize()
    print("H2D 24.9 MB ms:", a.elapsed_time(b), "GB/s:", x.numel() * 4 / a.elapsed_time(b) / 1e6)
z = torch.zeros(1, device="cuda")
t = time.perf_counter()
for i in range(100): z.to("cpu")
print("D2H 4B sync us:", (time.perf_counter() - t) / 100 * 1e6)
