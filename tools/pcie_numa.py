"""Does the e2e leg's drop at N > 1 come from where the pinned blocks live?  Times concurrent H2D + D2H of
one batch-sized block (268 MB each way) on the visible GPU, optionally after binding the process to the CPUs
of the GPU's NUMA node (first-touch then places the pinned pages on that node).

    python tools/pcie_numa.py DEVICE [bind]      # prints one line: device, node, bound?, ms, GB/s per direction
"""
import os
import sys
import time

import torch

dev_i = int(sys.argv[1])
bind = len(sys.argv) > 2 and sys.argv[2] == "bind"
p = torch.cuda.get_device_properties(dev_i)
bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
node = -1
try:
    node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
except OSError:
    pass
cpus = ""
if node >= 0:
    try:
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except OSError:
        pass
bound = False
if bind and cpus:
    want = set()
    for part in cpus.split(","):
        lo, _, hi = part.partition("-")
        want.update(range(int(lo), int(hi or lo) + 1))
    allowed = os.sched_getaffinity(0) & want
    if allowed:
        os.sched_setaffinity(0, allowed)
        bound = True

torch.cuda.set_device(dev_i)
n = 268435456 // 4
h_in, h_out = torch.zeros(n).pin_memory(), torch.zeros(n).pin_memory()
d_in, d_out = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
# start together with the other processes of the experiment (wall-clock slot)
slot = float(os.environ.get("START_AT", "0"))
while time.time() < slot:
    time.sleep(0.001)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


both(); torch.cuda.synchronize()
reps = 40
t0 = time.perf_counter()
for _ in range(reps):
    both()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / reps
print(f"gpu{dev_i} bus={bus} node={node} cpus={cpus} affinity={len(os.sched_getaffinity(0))} bound={bound} "
      f"both={ms:.2f} ms = {n * 4 / ms / 1e6:.1f} GB/s per direction", flush=True)
