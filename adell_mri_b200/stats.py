"""Device-side intensity statistics and scalers (K2/K3 + the exact intensity program).

Mirrors what the reference computes on the CPU inside its cached pre-transforms:
``ScaleIntensityd(minv=0, maxv=1)``, ``ConditionalRescalingd(500, 0.001)``, ``Offsetd(None)``,
``ScaleIntensityd(factor=-2/3)`` (/root/reference/adell_mri/transform_factory/transforms.py:
143-155,430-443,772-786; /root/reference/adell_mri/utils/monai_transforms/
image_intensity_ops.py:71-74,119-121) and the percentile scaler the north star names
(monai ``ScaleIntensityRangePercentilesd`` — ``np.percentile`` 'linear' semantics).
Everything stays on the device and on the current stream: no host round trip between the
statistics and the kernels that consume them.
"""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .plan import _TORCH_TO_ADELL

HIST_BITS = 11


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check_vols(vols: Sequence[torch.Tensor]):
    if len(vols) == 0:
        raise ValueError("no volumes")
    dev = vols[0].device
    if dev.type != "cuda":
        raise RuntimeError("adell_mri_b200.stats runs on CUDA devices only (no CPU fallback)")
    for v in vols:
        if v.device != dev or not v.is_contiguous() or v.dtype not in _TORCH_TO_ADELL:
            raise ValueError("volumes must be contiguous f32/i16/u8 tensors on one CUDA device")
    return dev


def vol_descriptors(vols: Sequence[torch.Tensor]) -> tuple[torch.Tensor, int]:
    """Upload the ``adell_vol`` descriptor array; returns (device buffer, max element count)."""
    dev = _check_vols(vols)
    arr = np.zeros(len(vols), np.dtype(_lib.Vol))
    arr["data"] = [v.data_ptr() for v in vols]
    arr["n"] = [v.numel() for v in vols]
    arr["dtype"] = [_TORCH_TO_ADELL[v.dtype] for v in vols]
    host = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).pin_memory()
    return host.to(dev, non_blocking=True), int(arr["n"].max())


def minmax(vols: Sequence[torch.Tensor], desc=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """``[n, 2]`` fp32 (min, max) per volume.  With ``desc`` (the uploaded descriptors of these very volumes) the
    per-volume checks are skipped; ``out``: a preallocated contiguous ``[n, 2]`` fp32 result."""
    if desc is None:
        dev = _check_vols(vols)
        d, max_n = vol_descriptors(vols)
    else:
        d, max_n = desc
        dev = d.device
    if out is None:
        out = torch.empty(len(vols), 2, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().adell_minmax(d.data_ptr(), len(vols), max_n, out.data_ptr(), _stream(dev)), "adell_minmax")
    return out


def meanstd(vols: Sequence[torch.Tensor], nonzero: bool = False, desc=None, raw_std: bool = False) -> torch.Tensor:
    """``[n, 2]`` fp32 (mean, population std) per volume.  A zero std is reported as 1 (monai
    NormalizeIntensity) unless ``raw_std`` (monai StdShiftIntensity: ``offset = factor * std``)."""
    dev = _check_vols(vols)
    d, max_n = desc if desc is not None else vol_descriptors(vols)
    out = torch.empty(len(vols), 2, dtype=torch.float32, device=dev)
    acc = torch.empty(len(vols), 3, dtype=torch.float64, device=dev)
    flags = (1 if nonzero else 0) | (2 if raw_std else 0)
    _lib.check(_lib.load().adell_meanstd(d.data_ptr(), len(vols), max_n, flags, acc.data_ptr(), out.data_ptr(),
                                         _stream(dev)), "adell_meanstd")
    return out


LABEL_COMBINE = {None: 0, "none": 0, "any": 1, "majority": 2}
LABEL_OP = {None: 0, "none": 0, "binary": 1, "cat": 2}


def label_map(vols: Sequence[torch.Tensor], combine: str | None, op: str | None, table: Sequence[float]) -> torch.Tensor:
    """Voxel-wise combination of up to 8 label maps (``"any"`` / ``"majority"``; ``None`` for a single
    map) followed by the label operator (``"binary"``: 1 where the value is in ``table``; ``"cat"``:
    index of the value in ``table``, else 0).  Returns a new fp32 tensor of the maps' shape."""
    dev = _check_vols(vols)
    if any(v.shape != vols[0].shape for v in vols):
        raise ValueError("label maps must share their shape")
    n_src = len(vols)
    ptrs = (C.c_void_p * n_src)(*[v.data_ptr() for v in vols])
    dts = (C.c_int32 * n_src)(*[_TORCH_TO_ADELL[v.dtype] for v in vols])
    tab = (C.c_float * max(len(table), 1))(*[float(x) for x in table])
    out = torch.empty(vols[0].shape, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().adell_label_map(ptrs, dts, n_src, LABEL_COMBINE[combine], LABEL_OP[op], tab, len(table),
                                           out.data_ptr(), out.numel(), _stream(dev)), "adell_label_map")
    return out


def mask_bbox(vols: Sequence[torch.Tensor]) -> torch.Tensor:
    """``[n, 6]`` int32 ``{lo0, hi0, lo1, hi1, lo2, hi2}`` (hi exclusive) of the non-zero voxels of each
    contiguous ``[S0, S1, S2]`` volume; an empty mask gives ``lo = INT32_MAX, hi = 0``."""
    dev = _check_vols(vols)
    if any(v.dim() != 3 for v in vols):
        raise ValueError("mask_bbox expects [S0, S1, S2] volumes")
    d, max_n = vol_descriptors(vols)
    shapes = torch.tensor([list(v.shape) for v in vols], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
    out = torch.empty(len(vols), 6, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().adell_mask_bbox(d.data_ptr(), shapes.data_ptr(), len(vols), max_n, out.data_ptr(), _stream(dev)),
               "adell_mask_bbox")
    return out


def gamma_map(vols: Sequence[torch.Tensor], minmax_dev: torch.Tensor, gammas, desc=None) -> list[torch.Tensor]:
    """monai AdjustContrast per volume: ``((x - min) / (range + 1e-7)) ** gamma * range + min`` with
    ``{min, max}`` read from ``minmax_dev`` (``[n, 2]`` fp32 on the device); returns new fp32 volumes."""
    dev = _check_vols(vols)
    d, max_n = desc if desc is not None else vol_descriptors(vols)
    outs = [torch.empty(v.shape, dtype=torch.float32, device=dev) for v in vols]
    ptrs = torch.tensor([o.data_ptr() for o in outs], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
    g = torch.tensor(np.broadcast_to(np.asarray(gammas, np.float32), (len(vols),)).copy()).pin_memory().to(dev, non_blocking=True)
    if minmax_dev.shape != (len(vols), 2) or minmax_dev.dtype != torch.float32 or not minmax_dev.is_contiguous():
        raise ValueError("minmax_dev must be a contiguous [n, 2] float32 tensor")
    _lib.check(_lib.load().adell_gamma_map(d.data_ptr(), ptrs.data_ptr(), minmax_dev.data_ptr(), g.data_ptr(), len(vols), max_n,
                                           _stream(dev)), "adell_gamma_map")
    return outs


def rician_map(x: torch.Tensor, noise1: torch.Tensor, noise2: torch.Tensor) -> torch.Tensor:
    """monai RandRicianNoise: ``sqrt((x + noise1)**2 + noise2**2)`` in fp32, one streaming device pass
    (``adell_rician_map``); returns a new tensor of ``x``'s shape."""
    if noise1.shape != x.shape or noise2.shape != x.shape:
        raise ValueError("rician_map: the noise volumes must have the shape of the input")
    if any(t.dtype != torch.float32 for t in (x, noise1, noise2)):
        raise ValueError("rician_map expects float32 tensors")
    x, noise1, noise2 = x.contiguous(), noise1.contiguous(), noise2.contiguous()
    dev = _check_vols([x, noise1, noise2])
    out = torch.empty_like(x)
    _lib.check(_lib.load().adell_rician_map(x.data_ptr(), noise1.data_ptr(), noise2.data_ptr(), out.data_ptr(), x.numel(),
                                            _stream(dev)), "adell_rician_map")
    return out


def resize(vols: Sequence[torch.Tensor], out_shape: Sequence[int], mode: str = "area") -> list[torch.Tensor]:
    """``F.interpolate(mode="area" | "nearest")`` of each contiguous fp32 ``[I0, I1, I2]`` volume to
    ``out_shape`` (``adell_resize``; area = ATen adaptive_avg_pool3d op for op, bit-identical)."""
    if mode not in ("area", "nearest"):
        raise NotImplementedError(f"resize mode '{mode}' is not on the device path (area / nearest are)")
    vols = [v.contiguous() for v in vols]
    dev = _check_vols(vols)
    if any(v.dim() != 3 or v.dtype != torch.float32 for v in vols):
        raise ValueError("resize expects float32 [I0, I1, I2] volumes")
    out_shape = [int(x) for x in out_shape]
    n = len(vols)
    block = torch.empty((n, *out_shape), dtype=torch.float32, device=dev)   # one allocation for all outputs
    outs = list(block.unbind(0))
    # one staging upload: source pointers, destination pointers, input extents
    host = np.empty(5 * n, np.int64)
    host[:n] = [v.data_ptr() for v in vols]
    stride = block.stride(0) * 4
    host[n:2 * n] = block.data_ptr() + stride * np.arange(n, dtype=np.int64)
    host[2 * n:].view(np.int32)[:3 * n] = np.array([list(v.shape) for v in vols], np.int32).reshape(-1)
    staged = torch.from_numpy(host).pin_memory().to(dev, non_blocking=True)
    d_src, d_dst, d_shp = staged[:n], staged[n:2 * n], staged[2 * n:]
    oshape = (C.c_int32 * 3)(*out_shape)
    _lib.check(_lib.load().adell_resize(d_src.data_ptr(), d_shp.data_ptr(), d_dst.data_ptr(), n, oshape,
                                        0 if mode == "area" else 1, _stream(dev)), "adell_resize")
    return outs


def scaler_coefs(stats: torch.Tensor, scaler: int, p0: float, p1: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """``[n, 6]`` coefficients of ``y = ((x*m0 - a)/d)*m1*m2 + b`` for one of the reference scalers."""
    n = stats.shape[0]
    coefs = out if out is not None else torch.empty(n, 6, dtype=torch.float32, device=stats.device)
    _lib.check(
        _lib.load().adell_scaler_coefs(stats.data_ptr(), n, scaler, float(p0), float(p1), coefs.data_ptr(), _stream(stats.device)),
        "adell_scaler_coefs",
    )
    return coefs


def coefs_to_affine(coefs: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Collapse the exact program into the fused-mode ``{scale, offset}`` pair (``[n, 2]``; ``out``: a preallocated
    contiguous ``[n, 2]`` fp32 destination, e.g. the rows a prepared K1 step reads through ``pre_dev``)."""
    n = coefs.shape[0]
    if out is None:
        out = torch.empty(n, 2, dtype=torch.float32, device=coefs.device)
    elif out.shape != (n, 2) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous [n, 2] float32 tensor")
    _lib.check(_lib.load().adell_coefs_to_affine(coefs.data_ptr(), n, out.data_ptr(), _stream(coefs.device)), "adell_coefs_to_affine")
    return out


def intensity_map(vols: Sequence[torch.Tensor], coefs: torch.Tensor, clip=None, desc=None) -> list[torch.Tensor]:
    """Exact (bit-faithful op order) elementwise scaler; returns new fp32 volumes."""
    dev = _check_vols(vols)
    d, max_n = desc if desc is not None else vol_descriptors(vols)
    outs = [torch.empty(v.shape, dtype=torch.float32, device=dev) for v in vols]
    ptrs = torch.tensor([o.data_ptr() for o in outs], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
    lo, hi = (0.0, 0.0) if clip is None else clip
    _lib.check(
        _lib.load().adell_intensity_map(
            d.data_ptr(), ptrs.data_ptr(), coefs.data_ptr(), len(vols), max_n, int(clip is not None), lo, hi, _stream(dev)
        ),
        "adell_intensity_map",
    )
    return outs


def _pass_schedule(dtype: torch.dtype):
    """(shift, bits) per radix pass over the 32-bit order-preserving key."""
    if dtype == torch.float32:
        return [(21, 11), (10, 11), (0, 10)]
    if dtype == torch.int16:
        return [(21, 11), (16, 5)]
    return [(24, 8)]


def numpy_virtual_index(n: int, q: float):
    """numpy's 'linear' quantile method in float64: virtual index ``(n - 1) * q`` (numpy evaluates exactly this
    expression for 'linear', not the mathematically equivalent ``_compute_virtual_index(n, q, 1, 1)``: the two
    round differently, and at a near-tie of the final float32 rounding the percentile moves by one ulp — found
    by tools/fuzz_stats.py), then ``_get_indexes`` / ``_get_gamma``: returns (lo, hi, gamma)."""
    quant = np.true_divide(np.float64(q), 100.0)
    vi = (n - 1) * quant
    if vi >= n - 1:
        return n - 1, n - 1, 0.0
    if vi < 0:
        return 0, 0, 0.0
    lo = int(np.floor(vi))
    return lo, lo + 1, float(vi - lo)


class _CudaKernels:
    """The three device entry points of the radix selection (C ABI)."""

    def __init__(self, vols):
        self.dev = _check_vols(vols)
        self.lib = _lib.load()
        self.desc, self.max_n = vol_descriptors(vols)
        self.n_vols = len(vols)
        self.dtype = vols[0].dtype
        self.st = _stream(self.dev)
        self.rank_cache = {}
        self.workspace, self.brackets_of, self.reuse_brackets = None, None, True

    def zeros(self, n, dtype):
        return torch.zeros(n, dtype=dtype, device=self.dev)

    def upload(self, arr: np.ndarray):
        return torch.from_numpy(arr).pin_memory().to(self.dev, non_blocking=True)

    def hist_pass(self, n_sel, shared, prefix, shift, bits, bins):
        _lib.check(self.lib.adell_hist_pass(self.desc.data_ptr(), self.n_vols, self.max_n, n_sel, int(shared),
                                            prefix.data_ptr(), shift, bits, bins.data_ptr(), self.st), "adell_hist_pass")

    def hist_select(self, bins, n_hist, n_sel, shift, bits, prefix, rank):
        _lib.check(self.lib.adell_hist_select(bins.data_ptr(), n_hist, n_sel, shift, bits, prefix.data_ptr(),
                                              rank.data_ptr(), self.st), "adell_hist_select")

    POOL_MAX_VOLS, POOL_MIN_N = 64, 4 * 32768   # what the pooled one-read path accepts (adell_quantile_keys)

    def quantile_keys(self, rank_dev, n_q, pooled_n: int = 0):
        """``adell_quantile_keys``: exact (lo, hi) keys of every (volume, quantile) — or of the POOLED data when
        ``pooled_n`` (the total element count) is given — in one full read."""
        need = self.lib.adell_quantile_workspace(self.n_vols, n_q, self.max_n, pooled_n)
        # the workspace belongs to this object (= these volumes): the brackets a call leaves in it are reused by the next
        # call with the same ranks instead of sampling the volumes again
        ws = self.workspace
        if ws is None or ws.numel() < need:
            ws = self.workspace = torch.empty(int(need) + 256, dtype=torch.uint8, device=self.dev)
            self.brackets_of = None
        reuse = self.reuse_brackets and self.brackets_of == (n_q, pooled_n, rank_dev.data_ptr())
        keys = torch.empty((1 if pooled_n else self.n_vols) * n_q * 2, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.adell_quantile_keys(self.desc.data_ptr(), self.n_vols, self.max_n, pooled_n, _TORCH_TO_ADELL[self.dtype], n_q,
                                                rank_dev.data_ptr(), keys.data_ptr(), ws.data_ptr(), ws.numel(), int(reuse), self.st),
                   "adell_quantile_keys")
        self.brackets_of = (n_q, pooled_n, rank_dev.data_ptr())
        self.last_workspace, self.last_pooled = ws, pooled_n
        return keys

    def fell_back(self, n_q) -> bool:
        """Whether the last :meth:`quantile_keys` call needed the radix fallback (synchronises; tests / diagnostics)."""
        out = C.c_int(0)
        _lib.check(self.lib.adell_quantile_fell_back(self.last_workspace.data_ptr(), self.n_vols, n_q, self.max_n, self.last_pooled,
                                                     C.byref(out)), "adell_quantile_fell_back")
        return bool(out.value)

    def finalize(self, prefix, frac, n_hist, n_q):
        out = torch.empty(n_hist, n_q, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.adell_percentile_finalize(prefix.data_ptr(), frac.data_ptr(), n_hist, n_q,
                                                      _TORCH_TO_ADELL[self.dtype], out.data_ptr(), self.st),
                   "adell_percentile_finalize")
        return out


def percentiles(
    vols: Sequence[torch.Tensor],
    qs: Sequence[float],
    dataset_wide: bool = False,
    all_reduce=None,
    total_n: int | None = None,
    kernels=None,
    one_read: bool = True,
) -> torch.Tensor:
    """Exact percentiles (numpy 'linear' method) of each volume: ``[n_vols, len(qs)]`` fp32.

    ``dataset_wide=True`` pools all volumes into one histogram (result ``[1, len(qs)]``);
    ``all_reduce(bins_tensor)`` — e.g. ``torch.distributed.all_reduce`` over NCCL — is then
    called on the int64 bin counts after every pass so that every rank selects identically,
    with ``total_n`` the element count over all ranks.  ``kernels`` is the device back end
    (the C ABI by default; the multi-rank host protocol is unit-tested with an injected one).
    Per-volume statistics take the one-read path (sampled brackets + exact selection among the few keys inside, radix
    passes as the in-call fallback); ``one_read=False`` forces the three radix passes.
    """
    kern = kernels if kernels is not None else _CudaKernels(vols)
    dtype = vols[0].dtype
    if any(v.dtype != dtype for v in vols):
        raise ValueError("percentiles: mixed dtypes")
    n_vols, n_q = len(vols), len(qs)
    n_sel = 2 * n_q
    if n_sel > 8:
        raise ValueError("at most 4 quantiles per call")
    n_hist = 1 if dataset_wide else n_vols
    # ranks / interpolation weights depend only on the element counts and the quantiles: computed and uploaded once per
    # (kernel object, quantiles) — a device-resident cache hands the same volumes back every epoch
    cache = getattr(kern, "rank_cache", None)
    ckey = (tuple(float(q) for q in qs), bool(dataset_wide), total_n)
    hit = cache.get(ckey) if cache is not None else None
    if hit is None:
        counts = [sum(v.numel() for v in vols) if total_n is None else total_n] if dataset_wide else [v.numel() for v in vols]
        ranks = np.zeros((n_hist, n_q, 2), np.uint64)
        frac = np.zeros((n_hist, n_q), np.float64)
        by_n = {}
        for h, n in enumerate(counts):
            row = by_n.get(n)
            if row is None:
                row = by_n[n] = [numpy_virtual_index(n, q) for q in qs]
            for j, (lo, hi, g) in enumerate(row):
                ranks[h, j] = (lo, hi)
                frac[h, j] = g
        hit = (kern.upload(ranks.reshape(-1).view(np.int64)), kern.upload(frac.reshape(-1)))
        if cache is not None:
            cache[ckey] = hit
    rank_dev, frac_dev = hit
    if all_reduce is None and isinstance(kern, _CudaKernels) and one_read:
        # statistics of one rank: one full read of every volume (adell_quantile_keys) instead of one per radix pass
        if not dataset_wide:
            return kern.finalize(kern.quantile_keys(rank_dev, n_q), frac_dev, n_hist, n_q)
        pooled = sum(v.numel() for v in vols)
        if total_n in (None, pooled) and len(vols) <= kern.POOL_MAX_VOLS and pooled >= kern.POOL_MIN_N:
            return kern.finalize(kern.quantile_keys(rank_dev, n_q, pooled), frac_dev, 1, n_q)
    if cache is not None:
        rank_dev = rank_dev.clone()   # adell_hist_select consumes the ranks
    prefix = kern.zeros(n_hist * n_sel, torch.int32)
    first = True
    for shift, bits in _pass_schedule(dtype):
        n_sel_eff = 1 if first else n_sel
        bins = kern.zeros((n_hist * n_sel_eff) << bits, torch.int64)
        kern.hist_pass(n_sel, dataset_wide, prefix, shift, bits, bins)
        if all_reduce is not None:
            all_reduce(bins)
        kern.hist_select(bins, n_hist, n_sel, shift, bits, prefix, rank_dev)
        first = False
    return kern.finalize(prefix, frac_dev, n_hist, n_q)
