"""Launch side of K1: uploads a composed :class:`~adell_mri_b200.plan.BatchPlan` and enqueues
the fused gather on the current CUDA stream (one launch per pass; a chain with a single
resample is exactly one launch for the whole batch).

Replaces the per-sample eager execution + ``safe_collate`` + pin + H2D of the reference's
DataLoader path (/root/reference/adell_mri/utils/utils.py:308-377,
/root/reference/adell_mri/entrypoints/segmentation/train.py:604-615).
"""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .plan import ITEM_DTYPE, BatchPlan

ISZ = _lib.ITEM_SIZE

#: number of K1 launches issued so far in this process (bench.py reports the delta per step)
launch_count = 0


class LaunchTimer:
    """Measurement aid: while installed (``engine.timer = LaunchTimer()``), every ``every``-th K1 launch is
    bracketed by a CUDA event pair recorded on the launching stream.  ``ms()`` (after a synchronize) returns
    the durations, tagged with the step label current at launch time (``timer.label = ...``), so that the
    passes of a multi-launch step can be summed."""

    def __init__(self, every: int = 1):
        self.every, self.n, self.pairs, self.label = max(int(every), 1), 0, [], None

    def begin(self, device):
        self.n += 1
        if (self.n - 1) % self.every:
            return None
        st = torch.cuda.current_stream(device)
        a = torch.cuda.Event(enable_timing=True)
        a.record(st)
        return a, st

    def end(self, tok):
        if tok is None:
            return
        b = torch.cuda.Event(enable_timing=True)
        b.record(tok[1])
        self.pairs.append((self.label, tok[0], b))

    def ms(self):
        return [(lab, a.elapsed_time(b)) for lab, a, b in self.pairs]


def aligned_bytes(n: int) -> np.ndarray:
    """Zeroed uint8 host buffer of ``n`` bytes starting on a 64-byte boundary: ``adell_item`` is declared 64-byte
    aligned (``include/adell_b200.h:107``) and the library refuses item buffers that are not
    (``ADELL_ERR_ALIGN``); numpy only promises 16 bytes for small allocations.  Pinned staging slots are page-aligned."""
    raw = np.zeros(int(n) + 64, np.uint8)
    o = (-raw.ctypes.data) % 64
    return raw[o: o + int(n)]


def aligned_items(n: int, like: np.ndarray | None = None) -> np.ndarray:
    """``n`` zeroed items (``ITEM_DTYPE``) on a 64-byte boundary, optionally filled from ``like``."""
    it = aligned_bytes(int(n) * ISZ).view(ITEM_DTYPE)
    if like is not None:
        it[:] = like
    return it


#: installed LaunchTimer or None
timer: LaunchTimer | None = None


class _PinnedRing:
    """Rotating pinned staging buffers for the parameter uploads.  torch's caching host allocator
    cannot hand a pinned block back while the copy that used it is still queued behind kernels,
    so under a full launch queue every upload would fall through to cudaHostAlloc (which blocks
    until the device is idle); a private ring with one event per slot avoids that."""

    def __init__(self, slots: int = 8):
        self.slots = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def acquire(self, n: int):
        """Next free slot with room for ``n`` bytes: ``(slot index, pinned uint8 numpy view [n])`` — the caller
        fills the view in place (the native composer writes the launch parameters straight into it)."""
        k = self.i
        self.i = (self.i + 1) % len(self.slots)
        if self.events[k] is not None:
            self.events[k].synchronize()  # normally long complete
        if self.slots[k] is None or self.slots[k].numel() < n:
            # cudaHostAlloc costs milliseconds: (re)size every slot in one go, with headroom, so a
            # steady-state loop never pins memory again
            cap = max(2 * n, 1 << 20)
            for j in range(len(self.slots)):
                if self.slots[j] is None or self.slots[j].numel() < cap:
                    if self.events[j] is not None:
                        self.events[j].synchronize()
                    self.slots[j] = torch.empty(cap, dtype=torch.uint8).pin_memory()
        return k, self.slots[k][:n].numpy()

    def upload(self, k: int, n: int, device: torch.device) -> torch.Tensor:
        # Blocks of a few hundred KB come out of the caching allocator's 2 MiB small-pool segments, two to six per
        # segment: a second live chunk of launch parameters then costs a cudaMalloc — 0.5 ms as a rule, 5-50 ms now and
        # then (seen in `bench.py`'s short runs: the whole queue drained behind it).  Asked for as MORE than 1 MiB (the
        # allocator's small / large boundary) they are carved out of one 20 MiB large-pool segment that the first upload
        # brought in.
        if n > (256 << 10):
            dev = torch.empty(max(n, (1 << 20) + 512), dtype=torch.uint8, device=device)[:n]
            dev.copy_(self.slots[k][:n], non_blocking=True)
        else:
            dev = self.slots[k][:n].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[k] = ev
        return dev

    def stage(self, buf: np.ndarray, device: torch.device) -> torch.Tensor:
        n = buf.shape[0]
        k, host = self.acquire(n)
        host[:] = buf
        return self.upload(k, n, device)


_rings: dict = {}


def _ring(device: torch.device) -> _PinnedRing:
    ring = _rings.get(device)
    if ring is None:
        ring = _rings[device] = _PinnedRing()
    return ring


def _stage(buf: np.ndarray, device: torch.device) -> torch.Tensor:
    return _ring(device).stage(buf, device)


def _require_cuda(device: torch.device):
    if device.type != "cuda":
        raise RuntimeError(
            "adell_mri_b200 executes on CUDA devices only (there is no CPU fallback); "
            f"got a plan living on {device}"
        )


def pack_launch(items: np.ndarray):
    """Host-side preparation of one launch: encodes the TMA descriptors of the eligible items
    (``adell_aug_prepare``) and packs ``items (768 B each) + int32 tile prefix (n+1)`` into one
    buffer.  Returns ``(uint8 buffer, n_items, LaunchInfo)``."""
    lib = _lib.load()
    n = items.shape[0]
    buf = aligned_bytes(n * ISZ + 4 * (n + 5))  # items, tile prefix (n + 1), chunk queues (4 words)
    it = buf[: n * ISZ].view(ITEM_DTYPE)
    it[:] = items
    tiles = buf[n * ISZ :].view(np.int32)
    info = _lib.LaunchInfo()
    _lib.check(lib.adell_aug_prepare(it.ctypes.data, n, tiles.ctypes.data, C.byref(info)), "adell_aug_prepare")
    return buf, n, info


def launch_packed(buf_dev: torch.Tensor, n: int, info, stream: int | None = None):
    """Enqueue one K1 launch from a device-resident packed buffer (see :func:`pack_launch`)."""
    global launch_count
    lib = _lib.load()
    if stream is None:
        stream = torch.cuda.current_stream(buf_dev.device).cuda_stream
    base = buf_dev.data_ptr()
    tok = timer.begin(buf_dev.device) if timer is not None else None
    _lib.check(lib.adell_aug_gather(base, base + n * ISZ, n, C.byref(info), C.c_void_p(stream)), "adell_aug_gather")
    if timer is not None:
        timer.end(tok)
    launch_count += lib.adell_aug_gather_launches()


def execute(plan: BatchPlan, dsts: Sequence[torch.Tensor]) -> None:
    """Run every recorded pass of ``plan``; volume ``i`` is written to ``dsts[i]`` (fp32
    ``[O0,O1,O2]`` views, any strides — typically channel slices of the collated batch)."""
    _require_cuda(plan.device)
    if len(dsts) != plan.n:
        raise ValueError("one destination per volume")
    shape = plan.shape
    dst_ptr = np.empty(plan.n, np.uint64)
    dst_stride = np.empty((plan.n, 3), np.int64)
    for i, d in enumerate(dsts):
        if d.dtype != torch.float32 or d.device != plan.device:
            raise ValueError("destinations must be float32 on the plan's device")
        if tuple(d.shape) != tuple(int(x) for x in shape[i]):
            raise ValueError(f"destination {i} has shape {tuple(d.shape)}, plan produces {tuple(shape[i])}")
        dst_ptr[i] = d.data_ptr()
        dst_stride[i] = d.stride()
    execute_ptrs(plan, dst_ptr, dst_stride)


_scratch: dict = {}


def _scratch_for(device: torch.device):
    """Allocator of the scratch volumes of multi-pass plans: ONE buffer per (device, stream), grown geometrically and
    reused by every later plan executed on that stream (launches of one stream are ordered, so the next plan's first
    pass cannot overtake the previous plan's last one).  A fresh ``torch.empty`` per plan made the caching allocator
    call cudaMalloc — a device-wide synchronisation — whenever a step closed more volumes than any step before it."""
    def alloc(n: int) -> torch.Tensor:
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        buf = _scratch.get(key)
        if buf is None or buf.numel() < n:
            buf = _scratch[key] = torch.empty(max(2 * n, 1 << 20), dtype=torch.float32, device=device)
        return buf[:max(n, 1)]
    return alloc


def execute_ptrs(plan: BatchPlan, dst_ptr: np.ndarray, dst_stride: np.ndarray, keep=None) -> None:
    """Like :func:`execute` with raw destination pointers (``[n]`` uint64, ``[n,3]`` element strides)."""
    _require_cuda(plan.device)
    if keep:
        plan.keep.extend(keep)
    with torch.cuda.device(plan.device):
        launches = plan.build_launches(dst_ptr, dst_stride, _scratch_for(plan.device))
        for items in launches:
            buf, n, info = pack_launch(items)
            dev = _stage(buf, plan.device)
            launch_packed(dev, n, info)
            plan.keep.append(dev)


class PreparedSteps:
    """Several steps composed and uploaded at once (one H2D copy), launched one by one."""

    def __init__(self, dev: torch.Tensor, offsets, sizes, infos_arr, keep):
        self.keep = [dev, infos_arr] + list(keep or [])
        self._lib = _lib.load()
        self._device = dev.device
        base = dev.data_ptr()
        self._args = [(base + int(o), base + int(o) + n * ISZ, n, C.byref(infos_arr[i])) for i, (o, n) in enumerate(zip(offsets, sizes))]
        self._per_launch = self._lib.adell_aug_gather_launches()
        self._gather = self._lib.adell_aug_gather

    def __len__(self):
        return len(self._args)

    def run(self, k: int, stream: int | None = None) -> None:
        """Enqueue step ``k`` on ``stream`` (a raw ``cudaStream_t``; default: torch's current stream of the device)."""
        global launch_count
        a = self._args[k]
        if stream is None:
            stream = torch.cuda.current_stream(self._device).cuda_stream
        tok = timer.begin(self._device) if timer is not None else None
        st = self._gather(a[0], a[1], a[2], a[3], C.c_void_p(stream))
        if timer is not None:
            timer.end(tok)
        if st != 0:
            _lib.check(st, "adell_aug_gather")
        launch_count += self._per_launch


def prepare_steps(plan: BatchPlan, dst_ptr: np.ndarray, dst_stride: np.ndarray, step_sizes, keep=None) -> PreparedSteps:
    """Compose once, launch many: ``plan`` holds the volumes of several consecutive steps
    (``step_sizes[k]`` volumes each, in order).  Host composition, TMA descriptor encoding and
    the parameter upload are done for all steps in one go, which amortises the Python / numpy
    overhead that would otherwise dominate a ~0.2 ms kernel.  Only single-pass plans (at most
    one resample per volume) can be split this way."""
    _require_cuda(plan.device)
    if plan.passes:
        raise ValueError("prepare_steps needs a single-pass plan (one resample per volume)")
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1]
    lib = _lib.load()
    sizes = [int(x) for x in step_sizes]
    if sum(sizes) != items.shape[0]:
        raise ValueError("step_sizes must add up to the number of volumes")
    # layout per step: items (768 B each) + int32 prefix + 4 chunk-queue words, padded to 128 B so every slice stays aligned
    ns = np.asarray(sizes, np.int64)
    step_bytes = ns * ISZ + ((4 * (ns + 5) + 127) // 128) * 128
    offs = np.concatenate([[0], np.cumsum(step_bytes)[:-1]]).astype(np.int64)
    total = int(step_bytes.sum())
    buf = aligned_bytes(total)
    if len(sizes) and (ns == ns[0]).all():
        # equal steps: one strided assignment places every step's items
        n0, stride = int(ns[0]), int(step_bytes[0])
        buf.reshape(len(sizes), stride)[:, : n0 * ISZ] = items.view(np.uint8).reshape(len(sizes), n0 * ISZ)
    else:
        start = 0
        for n, o in zip(sizes, offs):
            buf[o : o + n * ISZ] = items[start : start + n].view(np.uint8)
            start += n
    infos_arr = (_lib.LaunchInfo * len(sizes))()
    n32 = ns.astype(np.int32)
    tile_off = offs + ns * ISZ
    _lib.check(lib.adell_aug_prepare_steps(buf.ctypes.data, len(sizes), n32.ctypes.data, offs.ctypes.data,
                                           tile_off.ctypes.data, infos_arr), "adell_aug_prepare_steps")
    with torch.cuda.device(plan.device):
        dev = _stage(buf, plan.device)
    return PreparedSteps(dev, offs, sizes, infos_arr, [plan] + list(keep or []))


CHAIN_DTYPE = np.dtype(_lib.Chain)
_layout_cache: dict = {}


def _steps_layout(sizes: tuple):
    """Byte layout of several steps in one buffer: per step its items (768 B each), the int32 tile prefix and
    the chunk-queue words, padded to 128 B so every slice stays aligned.  Cached per tuple of step sizes."""
    hit = _layout_cache.get(sizes)
    if hit is None:
        ns = np.asarray(sizes, np.int64)
        step_bytes = ns * ISZ + ((4 * (ns + 5) + 127) // 128) * 128
        offs = np.concatenate([[0], np.cumsum(step_bytes)[:-1]]).astype(np.int64)
        if len(_layout_cache) > 256:
            _layout_cache.clear()
        hit = _layout_cache[sizes] = (ns.astype(np.int32), offs, offs + ns * ISZ, int(step_bytes.sum()))
    return hit


def prepare_chain_steps(chains: np.ndarray, step_sizes, device: torch.device, keep=None) -> PreparedSteps:
    """Native route of :func:`prepare_steps` for single-pass chains: ``chains`` (``CHAIN_DTYPE``, one per volume,
    the volumes of consecutive steps in order) are composed into items, prepared (tile policy, tensor maps) and
    written straight into a pinned staging slot by ONE call of ``adell_chain_prepare_steps`` — no numpy work per
    volume — then uploaded with one copy."""
    _require_cuda(device)
    if chains.dtype != CHAIN_DTYPE or not chains.flags.c_contiguous:
        raise ValueError("chains must be a contiguous array of adell_chain")
    sizes = tuple(int(x) for x in step_sizes)
    n32, offs, tile_off, total = _steps_layout(sizes)
    if int(n32.sum()) != chains.shape[0]:
        raise ValueError("step_sizes must add up to the number of chains")
    lib = _lib.load()
    infos_arr = (_lib.LaunchInfo * len(sizes))()
    with torch.cuda.device(device):
        ring = _ring(device)
        k, host = ring.acquire(total)
        _lib.check(lib.adell_chain_prepare_steps(chains.ctypes.data, host.ctypes.data, len(sizes), n32.ctypes.data,
                                                 offs.ctypes.data, tile_off.ctypes.data, infos_arr, 0), "adell_chain_prepare_steps")
        dev = ring.upload(k, total, device)
    return PreparedSteps(dev, offs, sizes, infos_arr, keep)


def compose_chains_host(chains: np.ndarray, step_sizes, plan_only: bool = True):
    """Host-only twin of :func:`prepare_chain_steps` (no device, no driver): returns ``(buffer, offsets, infos)``;
    used by the tests that compare the native composer with ``BatchPlan`` byte for byte."""
    sizes = tuple(int(x) for x in step_sizes)
    n32, offs, tile_off, total = _steps_layout(sizes)
    buf = aligned_bytes(total)
    infos_arr = (_lib.LaunchInfo * len(sizes))()
    _lib.check(_lib.load().adell_chain_prepare_steps(chains.ctypes.data, buf.ctypes.data, len(sizes), n32.ctypes.data,
                                                     offs.ctypes.data, tile_off.ctypes.data, infos_arr, int(plan_only)),
               "adell_chain_prepare_steps")
    return buf, offs, infos_arr


# ----------------------------------------------------------------------------- multi-pass sequences (native)
SEQ_DTYPE = np.dtype(_lib.Seq)
SEQ_OP_DTYPE = np.dtype(_lib.SeqOp)


class PreparedSeqSteps:
    """Steps whose volumes need several K1 passes (``adell_seq``): composed, prepared and uploaded at once; ``run(k)``
    enqueues the launches of step ``k`` in order (closed levels first, the final pass last)."""

    def __init__(self, dev: torch.Tensor, launches, n_launches: int, n_steps: int, keep):
        self.keep = [dev, launches] + list(keep or [])
        self._lib = _lib.load()
        self._device = dev.device
        base = dev.data_ptr()
        self._steps = [[] for _ in range(n_steps)]
        for i in range(n_launches):
            L = launches[i]
            self._steps[L.step].append((base + L.item_off, base + L.item_off + L.n_items * ISZ, L.n_items, C.byref(L.info)))
        self._per_launch = self._lib.adell_aug_gather_launches()
        self._gather = self._lib.adell_aug_gather

    def __len__(self):
        return len(self._steps)

    def launches(self, k: int) -> int:
        return len(self._steps[k])

    def run(self, k: int, stream: int | None = None) -> None:
        global launch_count
        if stream is None:
            stream = torch.cuda.current_stream(self._device).cuda_stream
        for a in self._steps[k]:
            tok = timer.begin(self._device) if timer is not None else None
            st = self._gather(a[0], a[1], a[2], a[3], C.c_void_p(stream))
            if timer is not None:
                timer.end(tok)
            if st != 0:
                _lib.check(st, "adell_aug_gather")
            launch_count += self._per_launch


def _seq_call(seqs: np.ndarray, sizes, scratch_ptr: int, scratch_elems: int, host_ptr: int, host_bytes: int, launches, mode: int):
    n32 = np.asarray(sizes, np.int32)
    nl, used, sused = C.c_int32(0), C.c_int64(0), C.c_int64(0)
    st = _lib.load().adell_seq_prepare_steps(seqs.ctypes.data, len(sizes), n32.ctypes.data, scratch_ptr, scratch_elems, host_ptr,
                                             host_bytes, launches, len(launches), C.byref(nl), C.byref(used), C.byref(sused), mode)
    return st, nl.value, used.value, sused.value


def _seq_bytes_bound(seqs: np.ndarray, sizes) -> int:
    """Upper bound of the staging bytes: a volume's first op never closes a pass, every later one at most once."""
    items = int(np.maximum(seqs["n_ops"].astype(np.int64), 1).sum())
    return items * ISZ + sum((_lib.SEQ_MAX_OPS + 1) * (4 * (n + 5) + 128) for n in sizes)


def prepare_seq_steps(seqs: np.ndarray, step_sizes, device: torch.device, keep=None) -> PreparedSeqSteps:
    """Native route for multi-pass chains: ``seqs`` (``SEQ_DTYPE``, one per volume, the volumes of consecutive steps in
    order) are composed pass by pass (BatchPlan's closing rules), every launch is prepared and written straight into a
    pinned staging slot by ONE call of ``adell_seq_prepare_steps``, then uploaded with one copy.  Scratch volumes live in
    the (device, stream) scratch buffer, reused by every step."""
    _require_cuda(device)
    if seqs.dtype != SEQ_DTYPE or not seqs.flags.c_contiguous:
        raise ValueError("seqs must be a contiguous array of adell_seq")
    sizes = tuple(int(x) for x in step_sizes)
    if sum(sizes) != seqs.shape[0]:
        raise ValueError("step_sizes must add up to the number of sequences")
    max_l = len(sizes) * (2 * _lib.SEQ_MAX_OPS + 1)
    launches = (_lib.SeqLaunch * max_l)()
    with torch.cuda.device(device):
        alloc = _scratch_for(device)
        alloc(1)
        scratch = _scratch[(device, torch.cuda.current_stream(device).cuda_stream)]   # the WHOLE (device, stream) buffer
        ring = _ring(device)
        need = _seq_bytes_bound(seqs, sizes)
        for _ in range(3):
            k, host = ring.acquire(need)
            st, nl, used, sused = _seq_call(seqs, sizes, scratch.data_ptr(), scratch.numel(), host.ctypes.data, need, launches, 0)
            if st == _lib.ERR_NO_SPACE and sused > scratch.numel():
                alloc(int(sused))   # grows the (device, stream) buffer
                scratch = _scratch[(device, torch.cuda.current_stream(device).cuda_stream)]
                continue
            _lib.check(st, "adell_seq_prepare_steps")
            break
        else:
            raise RuntimeError("adell_seq_prepare_steps: scratch / staging space could not be provided")
        dev = ring.upload(k, used, device)
    return PreparedSeqSteps(dev, launches, nl, len(sizes), [scratch] + list(keep or []))


def compose_seqs_host(seqs: np.ndarray, step_sizes, mode: int = 2, scratch_ptr: int = 0, scratch_elems: int = 1 << 60):
    """Host-only twin of :func:`prepare_seq_steps` (no device): returns ``[(step, items_array), ...]`` in execution
    order.  ``mode`` 2: the composed items as they are (what ``BatchPlan.build_launches`` returns), 1: after
    ``adell_aug_plan``."""
    sizes = tuple(int(x) for x in step_sizes)
    max_l = len(sizes) * (2 * _lib.SEQ_MAX_OPS + 1)
    launches = (_lib.SeqLaunch * max_l)()
    need = _seq_bytes_bound(seqs, sizes)
    buf = aligned_bytes(max(need, 1))
    st, nl, used, sused = _seq_call(seqs, sizes, scratch_ptr, scratch_elems, buf.ctypes.data, need, launches, mode)
    _lib.check(st, "adell_seq_prepare_steps")
    out = []
    for i in range(nl):
        L = launches[i]
        out.append((L.step, buf[L.item_off: L.item_off + L.n_items * ISZ].view(ITEM_DTYPE).copy()))
    return out, sused
