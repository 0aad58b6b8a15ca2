"""cProfile of one bench workload's step on the GPU box (measurement tooling): python tools/workload_host_profile.py seg_crop [steps]"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import torch
import bench_workloads as BW

name = sys.argv[1] if len(sys.argv) > 1 else "seg_crop"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cls = {c.name: c for c in BW.ALL}[name] if hasattr(BW, "ALL") else {c.name: c for c in vars(BW).values() if isinstance(c, type) and getattr(c, "name", "")}[name]
wl = cls(torch.device("cuda:0"), 0, 1, 20240)
for i in range(12):
    wl.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    wl.step(12 + i)
torch.cuda.synchronize()
print(f"{name}: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms per step (wall, {steps} steps)")
pr = cProfile.Profile(); pr.enable()
for i in range(steps):
    wl.step(40 + i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
