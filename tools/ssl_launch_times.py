"""Per-launch K1 times of the SSL (config C) workload step (measurement tooling): python tools/ssl_launch_times.py [ssl|ssl_fast]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench_workloads as BW
from adell_mri_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "ssl"
cls = {c.name: c for c in vars(BW).values() if isinstance(c, type) and getattr(c, "name", "")}[name]
wl = cls(torch.device("cuda:0"), 0, 1, 20240)
for i in range(16):
    wl.step(i)
torch.cuda.synchronize()
engine.timer = engine.LaunchTimer()
for i in range(16, 48):
    engine.timer.label = i
    wl.step(i)
torch.cuda.synchronize()
t, engine.timer = engine.timer, None
by = {}
for lab, ms in t.ms():
    by.setdefault(lab, []).append(ms)
arr = [v for v in by.values()]
n = max(len(v) for v in arr)
print("launches per step:", sorted({len(v) for v in arr}))
for j in range(n):
    xs = [v[j] for v in arr if len(v) > j]
    print(f"launch {j}: mean {np.mean(xs):.4f} ms (min {min(xs):.4f}, max {max(xs):.4f}) over {len(xs)} steps")
print("sum per step: mean %.4f ms" % np.mean([sum(v) for v in arr]))
# items / tiles of the launches of one step
p = wl._prep
for a in p._steps[0]:
    print("launch args:", [int(x) if isinstance(x, (int, np.integer)) else type(x).__name__ for x in a][:4], "tiles", a[3].total_tiles if hasattr(a[3], "total_tiles") else a[3].contents.total_tiles if hasattr(a[3], "contents") else "?")
