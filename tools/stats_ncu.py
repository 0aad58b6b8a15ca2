"""One call of the per-volume percentile path on config D's volumes (for `ncu -k regex:st_quantile_main`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adell_mri_b200 import stats
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
vols = [torch.empty((208, 208, 64), device=dev).log_normal_(5.0, 0.6, generator=g).reshape(-1) for _ in range(96)]
kern = stats._CudaKernels(vols)
for _ in range(3):
    stats.percentiles(vols, [0.5, 99.5], kernels=kern)
torch.cuda.synchronize()
