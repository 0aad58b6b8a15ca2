"""GPU, world_size 2 over NCCL (needs two devices; skipped otherwise): config E — dataset-wide
percentiles with the per-pass histogram counts all-reduced over NCCL, followed by the scale +
affine gather on every rank with the shared {scale, offset}."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from adell_mri_b200 import _lib, dist as adist, stats
        from adell_mri_b200.plan import BatchPlan
        from adell_mri_b200 import engine
        from oracle import monai_restated as M

        R = np.random.RandomState(99)
        shape = (96, 80, 40)
        all_vols = [torch.from_numpy(R.lognormal(0, 1, shape).astype(np.float32)) for _ in range(6)]
        mine = adist.shard_indices(len(all_vols))
        vols = [all_vols[i].to(dev) for i in mine]
        pct = adist.dataset_percentiles(vols, [1.0, 99.0])          # [1, 2], identical on every rank
        pooled = np.concatenate([v.numpy().reshape(-1) for v in all_vols])
        ref = np.percentile(pooled, np.asarray([1.0, 99.0])).astype(np.float32)
        # scale to [0,1] with the shared percentiles, then one affine gather per local volume
        coefs = stats.scaler_coefs(pct.expand(len(vols), 2).contiguous(), _lib.SCALER_RANGE, 0.0, 1.0)
        pre = stats.coefs_to_affine(coefs)
        A = M.compose_affine(rotate=[0.1, -0.05, 0.2], translate=[2.0, -1.0, 0.5])
        plan = BatchPlan(vols).intensity_from_device(pre).affine(A.numpy(), "bilinear", "zeros")
        outs = [torch.empty(shape, device=dev) for _ in vols]
        engine.execute(plan, outs)
        torch.cuda.synchronize()
        worst = 0.0
        for i, o in zip(mine, outs):
            scaled = M.scale_intensity_range(all_vols[i][None], float(ref[0]), float(ref[1]), 0.0, 1.0)
            want = M.affine_resample(scaled, A, "bilinear", "zeros")[0]
            worst = max(worst, float((o.cpu() - want).abs().max()) / float(want.abs().max()))
        q.put((rank, pct.cpu().numpy()[0].tolist(), ref.tolist(), worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_dataset_wide_percentile_and_gather_over_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == res[0][2]      # bit-identical on both ranks and equal to numpy on the pooled data
    assert max(r[3] for r in res) <= 1e-4           # stated fp32 tolerance for scale + trilinear
