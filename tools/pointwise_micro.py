"""Launch times of the pointwise statistics-side maps (CUDA events): `adell_rician_map` (16 B per voxel:
x, n1, n2 read + y written) and `adell_gamma_map` (8 B per voxel) on config C's two views
[64, 1, 128, 128, 32] x 2 and one 512x512x128 volume, next to the torch expressions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from adell_mri_b200 import stats

PEAK = 6541.1


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for shape in [(128, 1, 128, 128, 32), (1, 512, 512, 128)]:
    x = torch.rand(shape, device="cuda")
    n1, n2 = torch.randn(shape, device="cuda") * 0.1, torch.randn(shape, device="cuda") * 0.1
    n = x.numel()
    ms = timed(lambda: stats.rician_map(x, n1, n2))
    ms_t = timed(lambda: torch.sqrt((x + n1) ** 2 + n2 ** 2))
    print(f"rician {shape}: adell {ms:.4f} ms = {16 * n / ms / 1e6:.0f} GB/s ({16 * n / ms / 1e6 / PEAK:.2f} of peak); torch expression {ms_t:.4f} ms")
    flat = [x.reshape(-1)]
    mm = stats.minmax(flat)
    ms = timed(lambda: stats.gamma_map(flat, mm, 1.7))
    print(f"gamma  {shape}: adell {ms:.4f} ms = {8 * n / ms / 1e6:.0f} GB/s ({8 * n / ms / 1e6 / PEAK:.2f} of peak)")

# K5 adell_resize: config C sources through the scaled crop (random windows -> 160x160x40), and one large volume;
# wrapper (allocation + one staging upload + launch) and the launch alone
import ctypes as C
import math

import numpy as np

from adell_mri_b200 import _lib

lib = _lib.load()
for n, src, out in [(128, (120, 130, 30), (160, 160, 40)), (128, (200, 200, 48), (160, 160, 40)), (1, (512, 512, 128), (256, 256, 64))]:
    vols = [torch.rand(src, device="cuda") for _ in range(n)]
    block = torch.empty((n, *out), device="cuda")
    d_src = torch.tensor([v.data_ptr() for v in vols], dtype=torch.int64, device="cuda")
    d_dst = torch.tensor([block[i].data_ptr() for i in range(n)], dtype=torch.int64, device="cuda")
    d_shp = torch.tensor([list(src)] * n, dtype=torch.int32, device="cuda")
    oshape = (C.c_int32 * 3)(*out)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    b = 4 * n * (math.prod(src) + math.prod(out))
    for mode in ("area", "nearest"):
        ms = timed(lambda: stats.resize(vols, out, mode))
        mk = timed(lambda: lib.adell_resize(d_src.data_ptr(), d_shp.data_ptr(), d_dst.data_ptr(), n, oshape, 0 if mode == "area" else 1, st))
        print(f"resize {mode} {n} x {src} -> {out}: wrapper {ms:.4f} ms, launch alone {mk:.4f} ms = {b / mk / 1e6:.0f} GB/s of input + output bytes "
              f"({b / mk / 1e6 / PEAK:.2f} of peak)")
