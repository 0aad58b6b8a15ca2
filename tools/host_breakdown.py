"""Host-side cost breakdown of one 16-step chunk of the default bench workload (ms)."""
import sys, time
sys.path.insert(0, ".")
import ctypes as C
import numpy as np, torch
import bench
from adell_mri_b200 import _lib, engine
from adell_mri_b200.plan import ITEM_DTYPE

dev = torch.device("cuda:0")
cache = bench.make_cache("seg", 32, dev, 0)
aug = bench.make_augmenter("seg").set_random_state(1)
out = {"image": torch.empty((8, 3, 256, 256, 32), device=dev), "mask": torch.empty((8, 1, 256, 256, 32), device=dev)}
batches = [cache[(i % 4) * 8:(i % 4) * 8 + 8] for i in range(16)]
lib = _lib.load()
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
REPS = 10
for rep in range(REPS + 2):
    if rep == 2:
        T.clear()
    t = time.perf_counter(); samples = [s for b in batches for s in b]; plan = aug.plan(samples); tick("draw+compose", t)
    t = time.perf_counter()
    ptrs, strides = [], []
    for b in batches:
        p, st = aug._dst(out, len(b)); ptrs.append(p); strides.append(st)
    items = plan.build_launches(np.concatenate(ptrs), np.concatenate(strides), None)[-1]
    tick("fill items", t)
    t = time.perf_counter()
    n = 32
    buf = engine.aligned_bytes(16 * (n * ISZ + 256))
    infos = []
    for k in range(16):
        o = k * (n * ISZ + 256)
        it = buf[o:o + n * ISZ].view(ITEM_DTYPE); it[:] = items[k * n:(k + 1) * n]
        tiles = buf[o + n * ISZ:o + n * ISZ + 4 * (n + 5)].view(np.int32)
        info = _lib.LaunchInfo()
        lib.adell_aug_prepare(it.ctypes.data, n, tiles.ctypes.data, C.byref(info)); infos.append(info)
    tick("aug_prepare (tensor maps)", t)
    t = time.perf_counter(); d = engine._stage(buf, dev); tick("stage+upload", t)
    t = time.perf_counter()
    for k in range(16):
        o = k * (n * ISZ + 256)
        engine.launch_packed(d[o:o + n * ISZ + 4 * (n + 5)], n, infos[k])
    tick("16 launches", t)
    torch.cuda.synchronize()
for k, v in T.items():
    print(f"{k}: {v / REPS * 1e3:.3f} ms per 16-step chunk ({v / REPS / 16 * 1e3:.4f} per step)")
