"""K4 `adell_mixup` launch time on config B's collated image batch [8, 3, 256, 256, 32] (CUDA events),
next to the reference's torch expression on the same device (x*f + x[perm]*(1-f))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adell_mri_b200 import batch_preprocessing as B

PEAK = 6541.1
for shape in [(8, 3, 256, 256, 32), (32, 4, 192, 192, 48)]:
    x = torch.rand(shape, device="cuda")
    y = torch.zeros(shape[0], device="cuda")
    g = np.random.default_rng(0)
    n = x.numel()

    def ours():
        B.mixup(x, y, 0.4, g)

    def torch_ref():
        f = torch.as_tensor(g.beta(0.4, 0.4, shape[0]), dtype=x.dtype, device=x.device).reshape(-1, 1, 1, 1, 1)
        p = g.permutation(shape[0])
        return x * f + x[p] * (1.0 - f)

    import ctypes as C
    from adell_mri_b200 import _lib
    lib = _lib.load()
    f_dev = torch.rand(shape[0], device="cuda")
    p_dev = torch.randperm(shape[0], device="cuda").to(torch.int32)
    out = torch.empty_like(x)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def kernel_only():
        lib.adell_mixup(x.data_ptr(), out.data_ptr(), f_dev.data_ptr(), p_dev.data_ptr(), None, shape[0], n // shape[0], st)

    for name, fn in (("adell_mixup kernel alone", kernel_only), ("adell_mri_b200.batch_preprocessing.mixup (draws + upload + labels)", ours), ("torch expression", torch_ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        # algorithmic bytes: every element of the batch read once and written once (8 B); the second
        # operand x[perm] is another sample of the same batch and is not credited again
        gbs = 8.0 * n / (ms * 1e-3) / 1e9
        print(f"{shape} {name}: {ms:.4f} ms, {n / ms / 1e6:.1f} Gvox/s, {gbs:.0f} GB/s algorithmic (8 B/elem) = {gbs / PEAK:.3f} of measured peak")
