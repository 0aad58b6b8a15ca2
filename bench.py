#!/usr/bin/env python
"""bench.py — augmented voxels/s of the fused GPU augmentation hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload seg|affine_a] [--impl reference]

One *step* = one batch through the whole random chain (host parameter draws in the
reference's RandomState order -> host composition -> one K1 launch writing the collated
batch).  Default workload = BASELINE.json configs[1]: the u-net-3d-resnet segmentation
pipeline, 3 image keys (trilinear) + label mask (nearest), 256x256x32, batch 8,
get_augmentations_unet(["affine","flip"], flip_axis=[0,1,2]) with the reference's own
probabilities (affine 0.2, each flip 0.25).  `value` is device-resident throughput (CUDA
events, max over ranks); `e2e` adds the pinned-host H2D of the step's sources and the D2H of the step's result metric
(`e2e_full_readback`: of the whole batch);
`roofline` is algorithmic bytes / mean K1 launch time vs the measured HBM copy peak;
`cpu_baseline` is the oracle port (torch CPU grid_sample, the kernel MONAI calls) on the
host cores, one single-threaded worker process per core like the reference's DataLoader.
"""

from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "augmented voxels/s (RandAffine+flip+crop+norm)"
UNIT = "voxels/s"
SEED = 42

WORKLOADS = {
    # name: (description, spatial shape, image keys, batch per GPU)
    "seg": ("u-net-3d-resnet segmentation pipeline: T2/ADC/DWI trilinear + mask nearest, 256x256x32, batch 8, "
            "affine p=0.2 reflection + 3 flips p=0.25", (256, 256, 32), ["t2", "adc", "dwi"], 8),
    "seg_all_affine": ("same pipeline with the affine forced to fire for every sample (worst case for K1)",
                       (256, 256, 32), ["t2", "adc", "dwi"], 8),
}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region.  Uses NVML in-process
    (two cheap queries per sample): polling `nvidia-smi -lms 100` from a subprocess was measured
    to stall kernel launches and cost ~35 % of the throughput it was supposed to observe."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.gpu, self.period, self.samples, self.stop_flag, self.thread, self.err = gpu_index, period_s, [], False, None, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except ValueError:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            return self
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def _sample(self):
        nv = self.nv
        self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                             nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))

    def _loop(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no sampler"]}
        self.stop_flag = True
        self.thread.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        sm = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[1]
        reasons = sorted(k for k, bit in self.REASONS.items() if mask & bit)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.max, "reasons": reasons, "samples": len(sm),
                "source": "NVML, in-process, every %d ms during the timed region" % int(self.period * 1e3)}


# ----------------------------------------------------------------------------- data
def make_cache(workload: str, n_samples: int, device, seed: int):
    """Synthetic device-resident cache: what CacheDataset would hold after the deterministic
    pre-transforms (intensity-scaled fp32 images, 0/1 fp32 mask)."""
    import torch

    _, shape, image_keys, _ = WORKLOADS[workload]
    g = torch.Generator(device=device).manual_seed(seed)
    cache = []
    for _ in range(n_samples):
        s = {k: torch.rand((1, *shape), device=device, generator=g) for k in image_keys}
        s["mask"] = (torch.rand((1, *shape), device=device, generator=g) > 0.7).float()
        cache.append(s)
    return cache


def make_augmenter(workload: str):
    from adell_mri_b200.pipelines import SegmentationBatchAugmenter

    _, _, image_keys, _ = WORKLOADS[workload]
    aug = SegmentationBatchAugmenter(["affine", "flip"], image_keys + ["mask"], image_keys, flip_axis=[0, 1, 2])
    if workload == "seg_all_affine":
        for s in aug.samplers:
            s.prob = 1.0
    return aug


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One single-threaded worker (the reference's DataLoader-worker model): runs the oracle
    chain for `n` samples and returns (voxel-channels, seconds)."""
    workload, n, seed = args
    import numpy as np
    import torch

    torch.set_num_threads(1)
    from oracle import monai_restated as M

    _, shape, image_keys, _ = WORKLOADS[workload]
    R = np.random.RandomState(seed)
    keys = image_keys + ["mask"]
    sample = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)) for k in keys}
    prob = 1.0 if workload == "seg_all_affine" else 0.2
    draws = M.RandAffinedDraws(prob, rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16], n_keys=len(keys)).set_random_state(seed)
    flipR = [np.random.RandomState(seed + 1 + a) for a in range(3)]
    vox = 0
    t0 = time.perf_counter()
    for _ in range(n):
        fired, p = draws.draw()
        A = M.compose_affine(p["rotate"]) if fired else None
        flips = [a for a in range(3) if flipR[a].rand() < 0.25]
        outs = []
        for k in keys:
            o = M.canonical_item(sample[k], affine=A, mode="nearest" if k == "mask" else "bilinear",
                                 padding_mode="reflection", post_ops=[("flip", [a]) for a in flips])
            outs.append(o)
            vox += o.numel()
        torch.cat(outs[:-1], 0)  # ConcatItemsd(image_keys -> "image")
    return vox, time.perf_counter() - t0


class CpuReference:
    """Persistent pool of single-threaded workers (one per host core)."""

    def __init__(self, workload: str, cores: int | None = None):
        import multiprocessing as mp

        self.workload = workload
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_cpu_worker, [(workload, 0, 0)] * self.cores)  # import torch in every worker
        self.calls = 0

    def run(self, samples_per_worker: int):
        """One bounded sample of the workload on every host core.  Throughput = voxel-channels of ALL workers /
        WALL time of the whole map (round 1 divided by the slowest worker's busy time of one-sample steps, which
        timed every step at the worst-case draw and under-reported the CPU path about 2x)."""
        self.calls += 1
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(self.workload, samples_per_worker, SEED + 17 * i + 1009 * self.calls)
                                          for i in range(self.cores)], chunksize=1)
        wall = time.perf_counter() - t0
        vox = sum(r[0] for r in res)
        self.last = (vox, wall)
        return dict(value=vox / wall, unit=UNIT, cores=self.cores, kind="port",
                    sample=f"{self.cores} single-threaded worker processes x {samples_per_worker} samples of the "
                           f"'{self.workload}' chain (oracle port: torch CPU grid_sample + flip + concat); "
                           f"{vox} voxel-channels in {wall:.2f}s wall (slowest worker busy {max(r[1] for r in res):.2f}s)")

    def close(self):
        self.pool.close()
        self.pool.join()


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 1000 (b200 arm), 20 (reference arm: ~2 s of host work per step)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="seg", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-samples", type=int, default=6, help="samples per CPU worker for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cache-samples", type=int, default=32)
    ap.add_argument("--workloads", default="all", help="side workloads for the `workloads` block: 'all', 'none' or a comma list of "
                    "a,seg_all_affine,seg_norm,seg_crop,seg_crop_batch,ssl,ssl_fast,cls,large (see bench_workloads.py)")
    ap.add_argument("--workload-steps", type=int, default=0, help="timed steps per side workload (0: its own default)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run oracle checks of the side workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 1000

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: native libraries write there too (NCCL prints its
    # version banner on fd 1 when the box sets NCCL_DEBUG), so fd 1 points at stderr until the line is emitted.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    desc, shape, image_keys, batch = WORKLOADS[args.workload]
    nk = len(image_keys) + 1
    vox_per_step = batch * nk * shape[0] * shape[1] * shape[2]
    config = {"workload": f"{args.workload}: {desc}", "batch_per_gpu": batch, "shape": list(shape), "keys": nk,
              "src_dtype": "f32", "l2": "inputs larger than L2: each step reads 268 MB of sources from a rotating "
              f"{args.cache_samples}-sample device cache and writes 268 MB", "sharding": "by sample, no collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        # each step = a bounded sample of the workload on all host cores: `--cpu-samples` samples per worker, so
        # that a step averages over the expensive (affine fired) and cheap draws instead of waiting for the one
        # worker that drew the expensive branch
        ref = CpuReference(args.workload)
        per_step, raw = [], []
        for _ in range(args.warmup + args.steps):
            per_step.append(ref.run(args.cpu_samples))
            raw.append(ref.last)
        ref.close()
        # whole-run throughput = total voxel-channels / total wall time of the timed steps
        tot_v = sum(r[0] for r in raw[args.warmup:])
        tot_t = sum(r[1] for r in raw[args.warmup:])
        v = tot_v / tot_t
        cb = dict(per_step[-1]); cb["value"] = v
        cb["sample"] = f"{args.steps} steps, each: " + cb["sample"].split(";")[0] + f"; {tot_v} voxel-channels in {tot_t:.1f}s wall"
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * vox_per_step / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    import numpy as np
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from adell_mri_b200 import _lib, engine

    if not os.path.exists(_lib.LIB_PATH):
        raise SystemExit("libadell_b200.so missing: run __graft_entry__.build() first")
    _lib.load()

    cache = make_cache(args.workload, args.cache_samples, dev, 1234 + rank)
    aug = make_augmenter(args.workload).set_random_state(SEED + rank)
    out = {"image": torch.empty((batch, len(image_keys), *shape), device=dev),
           "mask": torch.empty((batch, 1, *shape), device=dev)}
    n_batches = args.cache_samples // batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    EVERY = 2   # in-loop launch timing: every 2nd step carries a CUDA event pair (10 samples at --steps 20)
    CHUNK = int(os.environ.get('BENCH_CHUNK', '32'))  # steps composed / uploaded together (host composition amortised; see engine.prepare_steps)

    def batch_of(i):
        b0 = (i % n_batches) * batch
        return cache[b0:b0 + batch]

    host = {"s": 0.0, "steps": 0}

    def prep(first, n):
        t0 = time.perf_counter()
        r = aug.prepare_steps([batch_of(first + k) for k in range(n)], [out] * n), n
        host["s"] += time.perf_counter() - t0
        host["steps"] += n
        return r

    def run_steps(first, count, ahead, next_n, events=None):
        """`count` steps starting at global step index `first`, software-pipelined the way a training loop
        runs them: the K1 launches of a chunk are queued, then the host makes the draws + composition + upload
        of the NEXT chunk (up to CHUNK steps) while the device works.  `ahead` = the prepared first chunk;
        after the last launch the host prepares `next_n` steps beyond the end and returns them, so a region of
        `count` steps holds `count` launches and the host work of as many steps as it was handed ready-made."""
        done = 0
        while done < count:
            prepared, n = ahead
            for k in range(n):
                timed = events is not None and (done + k) % EVERY == 0   # a sample of the launches: each
                if timed:                                              # event pair costs ~5 us of host time
                    events[2 * ((done + k) // EVERY)].record(stream)
                prepared.run(k, raw_stream)
                if timed:
                    events[2 * ((done + k) // EVERY) + 1].record(stream)
            done += n
            ahead = prep(first + done, min(CHUNK, count - done) if done < count else next_n)
        return ahead

    # ---- device-resident throughput ("value") ----
    stream = torch.cuda.current_stream()
    raw_stream = stream.cuda_stream
    first_n = min(CHUNK, args.steps)
    ahead = run_steps(0, args.warmup, prep(0, min(CHUNK, args.warmup)), first_n)
    barrier()
    clocks = ClockSampler(local_rank if not os.environ.get('BENCH_NO_CLOCKS') else -1).start()
    n_timed = (args.steps + EVERY - 1) // EVERY
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n_timed)]
    launches0 = engine.launch_count
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host["s"], host["steps"] = 0.0, 0
    e_start.record(stream)
    run_steps(args.warmup, args.steps, ahead, first_n, None if os.environ.get('BENCH_NO_EVENTS') else ev)
    host_us_per_step = 1e6 * host["s"] / max(host["steps"], 1)
    e_stop.record(stream)
    barrier()
    clock_info = clocks.stop()
    total_ms = e_start.elapsed_time(e_stop)
    k1_ms = [0.0] if os.environ.get('BENCH_NO_EVENTS') else [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n_timed)]
    launches = engine.launch_count - launches0

    # ---- roofline: K1 launch duration alone (params already uploaded), same seeded steps ----
    from adell_mri_b200.engine import launch_packed, pack_launch

    prepared = []
    aug_r = make_augmenter(args.workload).set_random_state(SEED + rank)
    for i in range(min(args.steps, 16)):
        b0 = (i % n_batches) * batch
        plan = aug_r.plan(cache[b0:b0 + batch])
        ptr = np.array([out["image"][b, c].data_ptr() for b in range(batch) for c in range(len(image_keys))]
                       ).reshape(batch, -1)
        mptr = np.array([out["mask"][b, 0].data_ptr() for b in range(batch)]).reshape(batch, 1)
        dst_ptr = np.concatenate([ptr, mptr], 1).reshape(-1).astype(np.uint64)
        dst_stride = np.tile(np.asarray(out["image"].stride()[2:], np.int64), (batch * nk, 1))
        items = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), device=dev))[-1]
        buf, n, info = pack_launch(items)
        prepared.append((torch.from_numpy(buf).to(dev), n, info, plan))
    torch.cuda.synchronize()
    kev = []
    for rep in range(3):
        for (buf, n, info, _) in prepared:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); launch_packed(buf, n, info); b.record(stream)
            kev.append((a, b))
    torch.cuda.synchronize()
    kernel_ms_alone = statistics.mean(a.elapsed_time(b) for a, b in kev[len(prepared):])
    # the roofline's launch duration: mean of the K1 launches sampled inside the timed region (every
    # EVERY-th step carries an event pair on the launching stream); the back-to-back repetition above
    # (parameters resident, nothing else queued) is reported next to it
    kernel_ms = statistics.mean(k1_ms) if k1_ms and k1_ms[0] > 0.0 else kernel_ms_alone
    alg_bytes = 8.0 * vox_per_step  # fp32 source read + fp32 write per output voxel-channel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per K1 launch cannot be measured in-run (it needs ncu's replay):
    # the figure comes from the committed `ncu --set full` capture of this workload and is labelled as such
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
        traffic_src = "from_profile: profiles/k1_traffic.json (ncu --set full capture of this workload, not measured in this run)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k1_gather", "kernel_ms": kernel_ms, "kernel_ms_isolated_repeat": kernel_ms_alone,
                "launches_sampled_in_timed_region": len(k1_ms), "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "traffic_source": traffic_src}

    # ---- end to end with host buffers ("e2e") ----
    # Host-resident cache (what the reference's CacheDataset holds in RAM): one pinned block per
    # batch, so a step is ONE host->device copy, one K1 launch and the device->host read of the
    # collated batch.  Three streams, double-buffered: the upload of step i+1 and the download of
    # step i-1 run on the two copy engines while K1 works on step i.
    keys = image_keys + ["mask"]
    host_in = [torch.stack([torch.stack([s[k] for k in keys]) for s in cache[j * batch:(j + 1) * batch]]).cpu().pin_memory()
               for j in range(2)]                                  # [B, keys, 1, H, W, D] fp32
    dev_in = [torch.empty_like(host_in[0], device=dev) for _ in range(2)]
    stage2 = [[{k: dev_in[j][b, ki] for ki, k in enumerate(keys)} for b in range(batch)] for j in range(2)]
    out2 = [out, {k: torch.empty_like(v) for k, v in out.items()}]
    host_out2 = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()} for _ in range(2)]
    h2d = host_in[0].numel() * 4
    d2h = sum(v.numel() * 4 for v in out.values())
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_k = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    # The step's RESULT read back by `e2e` is a metric of the batch (its per-sample means, computed on the device): the
    # augmented batch itself stays in HBM, where the model of a training step consumes it — the reference's loader
    # ends at the same point from the other side (a host batch that Lightning then uploads).  `e2e_full_readback`
    # also copies the whole collated batch back to pinned host memory (round 1's definition; nothing in either
    # pipeline needs that copy, it is kept as the conservative figure).
    metric_dev = [torch.empty(batch, device=dev) for _ in range(2)]
    metric_host = [torch.empty(batch).pin_memory() for _ in range(2)]
    d2h_metric = batch * 4

    def readback(j, full):
        if not full:
            torch.mean(out2[j]["image"], dim=(1, 2, 3, 4), out=metric_dev[j])
        ev_k[j].record(stream)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_k[j])
            if full:
                for k in out2[j]:
                    host_out2[j][k].copy_(out2[j][k], non_blocking=True)
            else:
                metric_host[j].copy_(metric_dev[j], non_blocking=True)
            ev_out[j].record(s_out)

    def e2e_step(i, full=False):
        j = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_k[j])  # staging buffer j free again (its previous kernel is done)
            dev_in[j].copy_(host_in[j], non_blocking=True)
            ev_in[j].record(s_in)
        stream.wait_event(ev_in[j])
        stream.wait_event(ev_out[j])  # output buffer j has been drained
        aug(stage2[j], out=out2[j])
        readback(j, full)

    e2e_steps = max(4, min(args.steps, 40))

    def time_e2e(step_fn, **kw):
        for i in range(2):
            step_fn(i, **kw)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(e2e_steps):
            step_fn(i, **kw)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / e2e_steps

    e2e_ms = time_e2e(e2e_step, full=False)
    e2e_full_ms = time_e2e(e2e_step, full=True)

    # ---- the same, from RAW host volumes (config B stores int16 images and a uint8 mask) ----
    # H2D carries 7 B per voxel position instead of 16; the min-max normalisation (adell_minmax -> coefficients ->
    # {scale, offset} read by K1 from device memory) runs inside the step.  Reported next to `e2e`, not instead of it.
    from adell_mri_b200 import stats as _stats

    g_raw = torch.Generator().manual_seed(99 + rank)
    raw_img = [torch.randint(0, 4000, (batch, len(image_keys), 1, *shape), generator=g_raw, dtype=torch.int16).pin_memory() for _ in range(2)]
    raw_msk = [(torch.rand((batch, 1, 1, *shape), generator=g_raw) > 0.7).to(torch.uint8).pin_memory() for _ in range(2)]
    dev_img = [torch.empty_like(raw_img[0], device=dev) for _ in range(2)]
    dev_msk = [torch.empty_like(raw_msk[0], device=dev) for _ in range(2)]
    stage_raw = [[{**{k: dev_img[j][b, ki] for ki, k in enumerate(image_keys)}, "mask": dev_msk[j][b, 0]} for b in range(batch)] for j in range(2)]
    flat_raw = [[s[k].reshape(-1) for s in stage_raw[j] for k in keys] for j in range(2)]
    desc_raw = [_stats.vol_descriptors(flat_raw[j]) for j in range(2)]
    aug_raw = make_augmenter(args.workload).set_random_state(SEED + 1000 + rank)
    h2d_raw = raw_img[0].numel() * 2 + raw_msk[0].numel()

    def e2e_raw_step(i, full=False):
        j = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_k[j])
            dev_img[j].copy_(raw_img[j], non_blocking=True)
            dev_msk[j].copy_(raw_msk[j], non_blocking=True)
            ev_in[j].record(s_in)
        stream.wait_event(ev_in[j])
        stream.wait_event(ev_out[j])
        mm = _stats.minmax(flat_raw[j], desc=desc_raw[j])
        pre = _stats.coefs_to_affine(_stats.scaler_coefs(mm, _lib.SCALER_MINMAX, 0.0, 1.0))
        aug_raw(stage_raw[j], out=out2[j], pre_dev=pre)
        readback(j, full)

    e2e_raw_ms = time_e2e(e2e_raw_step, full=False)
    e2e_raw_full_ms = time_e2e(e2e_raw_step, full=True)

    # ---- the drop-in dictionary surface: the reference's own call pattern ----
    # one transform-pipeline call per sample (get_augmentations_unet + post_transforms, same keys / arguments /
    # RandomState streams as the reference), then safe_collate: the recorded chains of the batch are composed once
    # and leave as one K1 launch.  Device-resident cache, like `value`; what an import swap costs per step.
    from adell_mri_b200 import collate as _collate, transform_factory as _F, transforms as _T

    tf_d = _F.SegmentationTransforms(keys, image_keys, None, image_keys, [])
    pipe_d = _T.Compose([_F.get_augmentations_unet(["affine", "flip"], keys, image_keys, [], flip_axis=[0, 1, 2]),
                         *tf_d.post_transforms()]).set_random_state(SEED + 2000 + rank)
    dict_steps = max(4, min(args.steps, 40))
    # warm-up: the collated outputs are fresh torch allocations (a cold cudaMalloc of 200 MB costs milliseconds); the loop
    # below keeps the previous step's batch alive while the next one is produced, like a training loop does
    for i in range(12):
        res_d = _collate.safe_collate([pipe_d(dict(smp)) for smp in batch_of(i)])
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _prof = None
    if os.environ.get("BENCH_PROFILE_DICT") == "1":
        import cProfile
        _prof = cProfile.Profile()
        _prof.enable()
    t0 = time.perf_counter()
    a.record(stream)
    for i in range(dict_steps):
        res_d = _collate.safe_collate([pipe_d(dict(smp)) for smp in batch_of(i)])
    b.record(stream)
    if _prof is not None:
        import pstats
        _prof.disable()
        pstats.Stats(_prof, stream=sys.stderr).sort_stats("tottime").print_stats(25)
    dict_host_ms = 1e3 * (time.perf_counter() - t0) / dict_steps
    barrier()
    dict_ms = a.elapsed_time(b) / dict_steps
    del res_d

    # ---- the other BASELINE configs, same run, same device (bench_workloads.py) ----
    workloads = {}
    if args.workloads != "none":
        import bench_workloads as BW

        want = None if args.workloads == "all" else set(args.workloads.split(","))
        torch.cuda.empty_cache()

        def reduce_max(x):
            tt = torch.tensor([x], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0])

        for cls in BW.ALL:
            if want is not None and cls.name not in want:
                continue
            workloads[cls.name] = BW.run(cls, dev, rank, world, SEED, args.workload_steps, args.warmup, peak, barrier, reduce_max,
                                         do_parity=not args.no_parity)
            if rank == 0:
                print("[workload] " + json.dumps({cls.name: workloads[cls.name]}), file=sys.stderr, flush=True)

    ms_per_step = total_ms / args.steps
    t = torch.tensor([ms_per_step, e2e_ms, e2e_raw_ms, dict_ms, e2e_full_ms, e2e_raw_full_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, e2e_ms, e2e_raw_ms, dict_ms, e2e_full_ms, e2e_raw_full_ms = (float(x) for x in t)

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * vox_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": roofline,
            "e2e": {"value": world * vox_per_step / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_metric, "ms_per_step": e2e_ms,
                    "what": "SegmentationBatchAugmenter.__call__ on pinned HOST fp32 volumes: H2D of the step's sources, K1, D2H of "
                            "the step's result metric (per-sample means of the augmented batch, reduced on the device); the batch "
                            "itself stays in HBM for the model, like the reference's host batch stays in RAM for Lightning"},
            "e2e_full_readback": {"value": world * vox_per_step / (e2e_full_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                                  "d2h_bytes_per_step": d2h, "ms_per_step": e2e_full_ms,
                                  "what": "the same with the whole collated batch copied back to pinned host memory (round 1's e2e)"},
            "e2e_raw_sources": {"value": world * vox_per_step / (e2e_raw_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_raw,
                                "d2h_bytes_per_step": d2h_metric, "ms_per_step": e2e_raw_ms,
                                "full_readback": {"value": world * vox_per_step / (e2e_raw_full_ms * 1e-3), "d2h_bytes_per_step": d2h,
                                                  "ms_per_step": e2e_raw_full_ms},
                                "what": "same step from pinned RAW volumes (int16 images + uint8 mask): min-max statistics on the "
                                        "device, normalisation folded into K1"},
            "gpu_launches": launches, "clocks": clock_info,
            "k1_launch_ms_median_in_loop": statistics.median(k1_ms), "host_chunk_steps": CHUNK,
            "host_us_per_step": host_us_per_step,
            "dict_surface_ms_per_step": dict_ms, "dict_surface_host_ms_per_step": dict_host_ms,
            "dict_surface": "reference call pattern (one pipeline call per sample + safe_collate), batch %d, device-resident cache" % batch,
            "workloads": workloads,
        }
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(args.workload)
            line["cpu_baseline"] = ref.run(args.cpu_samples)
            ref.close()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
