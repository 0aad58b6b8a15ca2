"""cProfile of the drop-in dictionary surface on the GPU box (measurement tooling): one pipeline call per sample + safe_collate,
the reference's call pattern (bench.py `dict_surface_ms_per_step`), batch 8 of config B."""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import torch
import bench
from adell_mri_b200 import collate as _collate, transform_factory as _F, transforms as _T

dev = torch.device("cuda:0")
cache = bench.make_cache("seg", 32, dev, 0)
keys = ["t2", "adc", "dwi", "mask"] if "t2" in cache[0] else list(cache[0].keys())
image_keys = keys[:-1]
tf_d = _F.SegmentationTransforms(keys, image_keys, None, image_keys, [])
pipe = _T.Compose([_F.get_augmentations_unet(["affine", "flip"], keys, image_keys, [], flip_axis=[0, 1, 2]), *tf_d.post_transforms()]).set_random_state(3)
def step(i):
    return _collate.safe_collate([pipe(dict(s)) for s in cache[(i % 4) * 8:(i % 4) * 8 + 8]])
for i in range(5): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(40): step(i)
torch.cuda.synchronize()
print("ms per step:", (time.perf_counter() - t0) / 40 * 1e3)
t0 = time.perf_counter()
for i in range(40): r = [pipe(dict(s)) for s in cache[(i % 4) * 8:(i % 4) * 8 + 8]]
print("record only, ms per step:", (time.perf_counter() - t0) / 40 * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(40): step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
