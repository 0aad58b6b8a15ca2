"""Per-step host and device times of one side workload (measurement tooling): python tools/workload_step_times.py cls [rank] [steps]
Shows which steps of a short driver-style run (5 warm-up + 10 timed) carry a composition burst, a fallback or an allocation."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench, bench_workloads as BW

name = sys.argv[1] if len(sys.argv) > 1 else "cls"
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 15
cls = {c.name: c for c in vars(BW).values() if isinstance(c, type) and getattr(c, "name", "")}[name]
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
wl = cls(dev, rank, 1, bench.SEED + 101 * rank)
parts = {}
def _wrap(obj, attr):
    fn = getattr(obj, attr, None)
    if fn is None:
        return
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            parts.setdefault(attr, []).append((time.perf_counter() - t) * 1e3)
    setattr(obj, attr, w)
for attr in ("_make", "_pre_dev"):
    _wrap(wl, attr)
from adell_mri_b200 import engine as _E, stats as _S
for mod, attr in ((_E._PinnedRing, "acquire"), (_E._PinnedRing, "upload")):
    _wrap(mod, attr)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
host = []
ev[0].record()
for i in range(steps):
    t = time.perf_counter()
    for v in parts.values():
        v.clear()
    wl.step(i)
    host.append(((time.perf_counter() - t) * 1e3, {k: (len(v), round(sum(v), 3), round(max(v), 3)) for k, v in parts.items() if v}))
    ev[i + 1].record()
torch.cuda.synchronize()
for i in range(steps):
    print(f"step {i:3d}: host {host[i][0]:8.3f} ms  device interval {ev[i].elapsed_time(ev[i + 1]):8.3f} ms  parts (calls, sum, max ms) {host[i][1]}")
