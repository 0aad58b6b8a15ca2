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
