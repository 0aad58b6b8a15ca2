"""CPU, world_size 2 over gloo: the N>1 host path — sample sharding and the dataset-wide
percentile protocol (histogram counts all-reduced after every radix pass).  The device entry
points are replaced by the numpy emulation in oracle/radix_select.py; the protocol, the rank /
bin bookkeeping and the collective calls are the product's (adell_mri_b200/dist.py, stats.py)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from adell_mri_b200 import dist as adist
        from oracle.radix_select import NumpyKernels

        R = np.random.RandomState(123)
        all_vols = [torch.from_numpy(R.lognormal(0, 1, (20, 16, 12)).astype(np.float32)) for _ in range(5)]
        all_vols[2][all_vols[2] < 0.7] = 0.0
        mine = adist.shard_indices(len(all_vols))
        vols = [all_vols[i] for i in mine]
        got = adist.dataset_percentiles(vols, [1.0, 50.0, 99.0], kernels=NumpyKernels(vols)).numpy()[0]
        pooled = np.concatenate([v.numpy().reshape(-1) for v in all_vols])
        ref = np.percentile(pooled, np.asarray([1.0, 50.0, 99.0])).astype(np.float32)
        q.put((rank, mine, got.tolist(), ref.tolist()))
    finally:
        dist.destroy_process_group()


def test_dataset_wide_percentile_over_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]          # disjoint shards covering everything
    assert res[0][2] == res[1][2] == res[0][3]                      # identical on both ranks, equal to numpy on the pooled data


def test_shard_indices_partition():
    from adell_mri_b200.dist import shard_indices

    for world in (1, 2, 4, 8):
        parts = [shard_indices(37, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(37))
