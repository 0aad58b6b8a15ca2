"""Measured PCIe ceiling of the box for the e2e leg of bench.py: pinned host <-> device copies of one
batch-sized block (268 MB), each direction alone and both at once on two streams."""
import torch

n = 268435456 // 4
h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    s = torch.cuda.current_stream(); s.wait_stream(s1); s.wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H concurrently", both)):
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms per 268 MB block = {0.268435456 / ms * 1e3:.1f} GB/s per direction")
