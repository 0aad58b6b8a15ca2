import cProfile, pstats, sys
sys.path.insert(0, ".")
import torch
import bench
dev = torch.device("cuda:0")
cache = bench.make_cache("seg", 32, dev, 0)
aug = bench.make_augmenter("seg").set_random_state(1)
out = {"image": torch.empty((8, 3, 256, 256, 32), device=dev), "mask": torch.empty((8, 1, 256, 256, 32), device=dev)}
batches = [cache[(i % 4) * 8:(i % 4) * 8 + 8] for i in range(16)]
for _ in range(3):
    p = aug.prepare_steps(batches, [out] * 16)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    p = aug.prepare_steps(batches, [out] * 16)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
