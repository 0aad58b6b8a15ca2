"""Host profile of the SSL (config C) workload's step on the GPU box (measurement tooling): where the ~0.67 ms of host
work per step of 64 two-view samples go (draws, native sequence composer, upload)."""
import sys, time
sys.path.insert(0, ".")
import torch, cProfile, pstats
import bench_workloads as BW
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
wl = BW.SSLTwoView(dev, 0, 1, 42)
for i in range(32): wl.step(i)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(32): wl.step(32 + i)
print("ssl host ms/step", (time.perf_counter() - t) / 32 * 1e3)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(64): wl.step(64 + i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
