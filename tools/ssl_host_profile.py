import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, cProfile, pstats
import bench_workloads as BW
dev=torch.device("cuda:0"); torch.cuda.set_device(0)
for cls in (BW.SSLTwoViewFast, BW.SSLTwoView):
    wl=cls(dev,0,1,42)
    for i in range(5): wl.step(i)
    torch.cuda.synchronize()
    t=time.perf_counter()
    for i in range(10): wl.step(5+i)
    print(cls.name, "host ms/step", (time.perf_counter()-t)/10*1e3)
    torch.cuda.synchronize()
    pr=cProfile.Profile(); pr.enable()
    for i in range(10): wl.step(15+i)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)
