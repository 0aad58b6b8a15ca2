import time, sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench
dev = torch.device("cuda:0")
w = sys.argv[1] if len(sys.argv) > 1 else "seg"
cache = bench.make_cache(w, 32, dev, 0)
aug = bench.make_augmenter(w).set_random_state(1)
out = {"image": torch.empty((8, 3, 256, 256, 32), device=dev), "mask": torch.empty((8, 1, 256, 256, 32), device=dev)}
batches = [cache[(i % 4) * 8:(i % 4) * 8 + 8] for i in range(16)]
for _ in range(2):
    p = aug.prepare_steps(batches, [out] * 16)
    for k in range(16): p.run(k)
torch.cuda.synchronize()
T0 = time.perf_counter()
tp = tl = 0.0
for c in range(6):
    t = time.perf_counter(); p = aug.prepare_steps(batches, [out] * 16); tp += time.perf_counter() - t
    t = time.perf_counter()
    for k in range(16): p.run(k)
    tl += time.perf_counter() - t
t = time.perf_counter(); torch.cuda.synchronize(); ts = time.perf_counter() - t
tot = time.perf_counter() - T0
print(f"{w}: total {tot*1e3:.2f} ms for 96 steps ({tot/96*1e3:.3f}/step); host prepare {tp*1e3:.2f}, host launch loop {tl*1e3:.2f}, final sync wait {ts*1e3:.2f}")
