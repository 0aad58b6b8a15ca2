import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from adell_mri_b200 import stats, _lib
dev=torch.device('cuda:0')
g=torch.Generator(device=dev).manual_seed(0)
for name,nv,shape in (("cls",96,(208,208,64)),("large",4,(512,512,128))):
    vols=[torch.empty(shape,device=dev).log_normal_(5.0,0.6,generator=g).reshape(-1) for _ in range(nv)]
    kern=stats._CudaKernels(vols)
    for one in (True,False):
        for _ in range(3): stats.percentiles(vols,[0.5,99.5],kernels=kern,one_read=one)
        torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): p=stats.percentiles(vols,[0.5,99.5],kernels=kern,one_read=one)
        b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b)/10
        nb=sum(v.numel() for v in vols)*4
        print(name,'one_read' if one else 'radix3',f'{ms:.3f} ms', f'{nb/ms/1e6:.0f} GB/s credited once', 'fell_back', kern.fell_back(2) if one else '-')
    if name == "large":
        for one in (True, False):
            for _ in range(3): stats.percentiles(vols,[1.0,99.0],dataset_wide=True,kernels=kern,one_read=one)
            torch.cuda.synchronize()
            a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): p=stats.percentiles(vols,[1.0,99.0],dataset_wide=True,kernels=kern,one_read=one)
            b.record(); torch.cuda.synchronize()
            ms=a.elapsed_time(b)/10
            print(name,'POOLED','one_read' if one else 'radix3',f'{ms:.3f} ms', f'{nb/ms/1e6:.0f} GB/s credited once', 'fell_back', kern.fell_back(2) if one else '-', p.cpu().numpy())
    del vols

# kernel-level split of one call (torch profiler, CUDA activities)
from torch.profiler import profile, ProfilerActivity
vols=[torch.empty((208,208,64),device=dev).log_normal_(5.0,0.6,generator=g).reshape(-1) for _ in range(96)]
kern=stats._CudaKernels(vols)
for _ in range(3): stats.percentiles(vols,[0.5,99.5],kernels=kern)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): stats.percentiles(vols,[0.5,99.5],kernels=kern)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
    print(f"{e.key[:70]:70s} n={e.count:3d} avg={e.device_time_total/e.count:8.1f} us")
