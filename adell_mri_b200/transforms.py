"""Dictionary transforms with the MONAI calling convention, executed lazily on the GPU.

The reference wires ``monai.transforms.*`` dictionary transforms into ``Compose`` pipelines
(/root/reference/adell_mri/transform_factory/augmentations.py:19-178,181-320,391-516;
/root/reference/adell_mri/transform_factory/transforms.py:53-67) that DataLoader workers run
eagerly on the CPU, one full-volume copy per op.  The classes here obey the same protocol
(callable ``dict -> dict | list[dict]``, ``.R`` / ``set_random_state`` / ``randomize`` on the
random ones, the same constructor arguments and the same RandomState draw order †) but do not
touch voxels: every op is *recorded* on a :class:`Pending` entry (a one-sample
:class:`~adell_mri_b200.plan.BatchPlan`).  :func:`adell_mri_b200.collate.safe_collate` then
concatenates the recorded chains of the whole batch and runs them as ONE fused K1 launch per
resample pass, writing straight into the collated ``[B, C, H, W, D]`` tensors.

† draw orders restated from MONAI 1.3-1.6 (not importable here): see oracle/monai_restated.py.
Transforms of the reference vocabulary that are not gathers / pointwise maps (blur, bias field,
Gibbs, grid distortion, low-res, coarse dropout ...) raise ``NotImplementedError`` — they are
outside the fused hot path (out of scope).
"""

from __future__ import annotations

import itertools
from typing import Callable, Hashable, Sequence

import numpy as np
import torch

from . import geometry
from .plan import BatchPlan
from .sampling import MAX_SEED, RandAffineSampler, rand_param

# --------------------------------------------------------------------------- execution mode
_MODE = {"strict": False, "fast": False, "noise": "injected"}
_EYE4 = np.eye(4, dtype=np.float32)


def set_mode(strict: bool | None = None, fast: bool | None = None, noise: str | None = None):
    """Global execution mode of newly created :class:`Pending` entries.

    ``strict``: bit-faithful ATen operation order for trilinear resampling (parity mode);
    ``fast``: consecutive resamples are composed into one matrix (documented deviation);
    ``noise``: ``"injected"`` (host RandomState.normal tensor, as the reference draws it) or
    ``"philox"`` (device-side counter-based normal with the host-drawn sigma)."""
    if strict is not None:
        _MODE["strict"] = bool(strict)
    if fast is not None:
        _MODE["fast"] = bool(fast)
    if noise is not None:
        if noise not in ("injected", "philox"):
            raise ValueError("noise must be 'injected' or 'philox'")
        _MODE["noise"] = noise
    return dict(_MODE)


class _Part:
    """Channels of one entry that share a recorded chain: the parent volumes, the chain as DATA — a list of
    ``(name, static, dyn, per_volume)`` — and the logical spatial shape the chain has reached.  Nothing numeric
    happens while a pipeline runs; :func:`replay_parts` turns the parts of a whole batch into BatchPlans, one per
    distinct chain signature, each op applied once to all the volumes that recorded it."""

    __slots__ = ("vols", "ops", "shape", "strict", "fast", "_sig", "meta")

    def __init__(self, vols, shape, strict, fast, ops=None, meta=None):
        self.vols, self.shape, self.strict, self.fast = vols, tuple(int(x) for x in shape), strict, fast
        self.ops = [] if ops is None else ops
        self._sig = None
        self.meta = meta   # what BatchPlan needs to know about the parents: gathered at replay (parent_meta)

    def copy(self):
        return _Part(self.vols, self.shape, self.strict, self.fast, list(self.ops), self.meta)

    def signature(self):
        if self._sig is None or self._sig[0] != len(self.ops):
            self._sig = (len(self.ops), (len(self.vols), self.strict, self.fast, self.vols[0].dtype, str(self.vols[0].device),
                                         tuple((o[0], o[1]) for o in self.ops)))
        return self._sig[1]


def parent_meta(part: _Part, source=None):
    """(ptr, stride, dtype, shape) arrays of a part's parent volumes.  Gathered once per part — and remembered on the
    SOURCE tensor the entry was made from (``source._adell_meta``): a device-resident cache hands the same tensor
    objects back every epoch, so after the first one this costs a dictionary lookup."""
    if part.meta is None:
        from .plan import _TORCH_TO_ADELL
        vols = part.vols
        for v in vols:
            if v.dim() != 3 or v.dtype not in _TORCH_TO_ADELL:
                raise TypeError("expected [H, W, D] float32 / int16 / uint8 volumes")
        part.meta = (np.array([v.data_ptr() for v in vols], np.uint64), np.array([v.stride() for v in vols], np.int64),
                     np.array([_TORCH_TO_ADELL[v.dtype] for v in vols], np.uint8), np.array([v.shape for v in vols], np.int64))
    return part.meta


def _shape_after(name, shape, dyn):
    """Logical spatial shape after one recorded op (the integer rules of BatchPlan.crop / center_crop / pad /
    spatial_pad / resample_to, on plain ints)."""
    if name == "crop":
        start, size = dyn
        out = []
        for c, st, sz in zip(shape, start, size):
            st = min(max(int(st), 0), c)
            out.append(min(int(sz), c - st))
        return tuple(out)
    if name == "center_crop":
        return tuple(c if int(r) <= 0 else min(int(r), c) for c, r in zip(shape, dyn[0]))
    if name == "pad":
        return tuple(c + int(b) + int(a) for c, b, a in zip(shape, dyn[0], dyn[1]))
    if name == "spatial_pad":
        return tuple(max(c, int(w)) if int(w) > 0 else c for c, w in zip(shape, dyn[0]))
    if name == "resample_to":
        return tuple(int(x) for x in dyn[1])
    return shape


class ChainRecorder:
    """``Pending.plan``: the methods of :class:`~adell_mri_b200.plan.BatchPlan` the dictionary transforms call,
    recording instead of composing (a tuple append and an integer shape update per call)."""

    __slots__ = ("parts",)

    def __init__(self, parts):
        self.parts = parts

    def _rec(self, name, static, dyn, per_volume=False):
        for p in self.parts:
            p.ops.append((name, static, dyn, per_volume))
            p.shape = _shape_after(name, p.shape, dyn)
        return self

    # integer ops
    def crop(self, start, size):
        return self._rec("crop", (), (tuple(int(x) for x in start), tuple(int(x) for x in size)))

    def crop_from_device(self, win_row: torch.Tensor, size):
        """Crop whose start is the int32[3] ``win_row`` in device memory (see BatchPlan.crop_from_device)."""
        size = tuple(int(x) for x in size)
        for p in self.parts:
            if p.ops:
                raise ValueError("a device-side crop window must be the first recorded op")
            p.ops.append(("crop_from_device", (), (int(win_row.data_ptr()), size, win_row), False))
            p.shape = tuple(min(c, s) for c, s in zip(p.shape, size))
        return self

    def center_crop(self, roi):
        return self._rec("center_crop", (), (tuple(int(x) for x in roi),))

    def pad(self, before, after):
        return self._rec("pad", (), (tuple(int(x) for x in before), tuple(int(x) for x in after)))

    def spatial_pad(self, spatial_size):
        return self._rec("spatial_pad", (), (tuple(int(x) for x in spatial_size),))

    def flip(self, mask):
        if type(mask) is not tuple:
            mask = tuple(bool(x) for x in np.asarray(mask).reshape(3))
        return self._rec("flip", (), (mask,))

    # resamples
    def affine(self, A, mode="bilinear", padding_mode="reflection", fired: bool = True):
        return self._rec("affine", (padding_mode,), (np.asarray(A, np.float32), mode, bool(fired)))

    def resample_to(self, step, out_size, mode="bilinear", padding_mode="border"):
        return self._rec("resample_to", (mode, padding_mode), (tuple(float(x) for x in step), tuple(int(x) for x in out_size)))

    # intensity / noise
    def intensity(self, scale=1.0, offset=0.0):
        return self._rec("intensity", (), (float(scale), float(offset)))

    def intensity_from_device(self, pre_dev: torch.Tensor):
        n = sum(len(p.vols) for p in self.parts)
        if pre_dev.shape != (n, 2) or pre_dev.dtype != torch.float32:
            raise ValueError("pre_dev must be a [n, 2] float32 tensor")
        o = 0
        for p in self.parts:
            if p.ops:
                raise ValueError("device-side intensity must be the first recorded op")
            p.ops.append(("intensity_from_device", (), (pre_dev[o:o + len(p.vols)],), True))
            o += len(p.vols)
        return self

    def add_philox_noise(self, std, seed, offset=0):
        o = 0
        for p in self.parts:
            off = np.broadcast_to(np.asarray(offset, np.uint64), (sum(len(q.vols) for q in self.parts),))[o:o + len(p.vols)]
            p.ops.append(("add_philox_noise", (), (np.float32(std), np.uint64(seed), off.copy()), True))
            o += len(p.vols)
        return self

    def add_noise(self, noise):
        o = 0
        for p in self.parts:
            p.ops.append(("add_noise", (), (list(noise[o:o + len(p.vols)]),), True))
            o += len(p.vols)
        return self

    @property
    def n(self) -> int:
        return sum(len(p.vols) for p in self.parts)

    @property
    def device(self):
        return self.parts[0].vols[0].device

    @property
    def shape(self) -> np.ndarray:
        return np.array([p.shape for p in self.parts for _ in p.vols], np.int64)

    def build(self) -> BatchPlan:
        """The composed plan of this entry alone."""
        return replay_parts(self.parts)


def replay_parts(parts: Sequence[_Part]) -> BatchPlan:
    """Compose the recorded chains of many parts (in order: the plan's volumes are the parts' volumes concatenated).
    Parts are grouped by chain signature (op names + static arguments, channel count, mode flags); every op of a
    group is ONE vectorised BatchPlan call over all of the group's volumes, its per-part arguments stacked."""
    groups: dict = {}
    for i, p in enumerate(parts):
        groups.setdefault(p.signature(), []).append(i)
    plans, order = [], []
    for sig, idx in groups.items():
        grp = [parts[i] for i in idx]
        nv = len(grp[0].vols)
        dev = grp[0].vols[0].device
        for p in grp:
            if p.vols[0].device != dev:
                raise ValueError("all parents of a plan must live on one device")
        metas = [parent_meta(p) for p in grp]
        cat = (lambda i: metas[0][i]) if len(grp) == 1 else (lambda i: np.concatenate([m[i] for m in metas]))
        plan = BatchPlan.from_arrays(cat(0), cat(1), cat(2), cat(3), dev, [p.vols for p in grp], fast=grp[0].fast, strict=grp[0].strict)
        rep = (lambda x: x) if nv == 1 else (lambda x: np.repeat(x, nv, axis=0))
        for j, (name, static, _, per_volume) in enumerate(grp[0].ops):
            dyn = [p.ops[j][2] for p in grp]
            if name == "crop":
                plan.crop(rep(np.array([d[0] for d in dyn], np.int64)), rep(np.array([d[1] for d in dyn], np.int64)))
            elif name == "crop_from_device":
                plan.crop_from_device(rep(np.array([d[0] for d in dyn], np.uint64)), rep(np.array([d[1] for d in dyn], np.int64)))
                plan.keep.extend(d[2] for d in dyn)
            elif name == "center_crop":
                plan.center_crop(rep(np.array([d[0] for d in dyn], np.int64)))
            elif name == "pad":
                plan.pad(rep(np.array([d[0] for d in dyn], np.int64)), rep(np.array([d[1] for d in dyn], np.int64)))
            elif name == "spatial_pad":
                plan.spatial_pad(rep(np.array([d[0] for d in dyn], np.int64)))
            elif name == "flip":
                plan.flip(rep(np.array([d[0] for d in dyn], bool)))
            elif name == "affine":
                fired = np.array([d[2] for d in dyn], bool)
                if fired.any():
                    modes = [d[1] for d in dyn for _ in range(nv)]
                    plan.affine(rep(np.stack([d[0] for d in dyn])), modes, static[0], where=None if fired.all() else rep(fired))
            elif name == "resample_to":
                plan.resample_to(rep(np.array([d[0] for d in dyn], np.float64)), rep(np.array([d[1] for d in dyn], np.int64)), static[0], static[1])
            elif name == "intensity":
                plan.intensity(scale=rep(np.array([d[0] for d in dyn], np.float64)), offset=rep(np.array([d[1] for d in dyn], np.float64)))
            elif name == "intensity_from_device":
                plan.intensity_from_device(dyn[0][0].contiguous() if len(dyn) == 1 else torch.cat([d[0] for d in dyn]).contiguous())
            elif name == "add_philox_noise":
                plan.add_philox_noise(rep(np.array([d[0] for d in dyn], np.float32)), rep(np.array([d[1] for d in dyn], np.uint64)),
                                      np.concatenate([d[2] for d in dyn]))
            elif name == "add_noise":
                plan.add_noise([t for d in dyn for t in d[0]])
            else:
                raise ValueError(name)
        plans.append(plan)
        order.extend((i, k) for i in idx for k in range(nv))
    if len(plans) == 1 and order == [(i, k) for i in range(len(parts)) for k in range(len(parts[0].vols))]:
        return plans[0]
    # volumes back into the callers' order (parts in order, their channels in order)
    plan = BatchPlan.concat(plans)
    want = {}
    pos = 0
    for i, p in enumerate(parts):
        for k in range(len(p.vols)):
            want[(i, k)] = pos
            pos += 1
    perm = np.empty(len(order), np.int64)     # perm[target position] = position in the concatenated plan
    for src, key in enumerate(order):
        perm[want[key]] = src
    return plan.permuted(perm)


class Pending:
    """One dict entry ``[C, H, W, D]`` whose transform chain is recorded, not yet executed."""

    def __init__(self, tensor: torch.Tensor | None = None, parts: list | None = None, meta: dict | None = None):
        if parts is None:
            if tensor.dim() != 4:
                raise ValueError("expected a channel-first [C, H, W, D] volume")
            cached = getattr(tensor, "_adell_part", None)   # (vols, meta) of an earlier entry made from this very tensor
            if cached is not None and cached[2] == tensor._version and cached[3] == tensor.data_ptr():
                part = _Part(cached[0], tensor.shape[1:], _MODE["strict"], _MODE["fast"], meta=cached[1])
            else:
                part = _Part([tensor[c] for c in range(tensor.shape[0])], tensor.shape[1:], _MODE["strict"], _MODE["fast"])
                try:
                    tensor._adell_part = (part.vols, parent_meta(part), tensor._version, tensor.data_ptr())
                except Exception:   # noqa: BLE001 — tensors that refuse attributes simply do not cache
                    pass
            parts = [part]
        self.parts = parts
        self.meta = meta if meta is not None else {}

    @property
    def plan(self) -> ChainRecorder:
        return ChainRecorder(self.parts)

    @property
    def n_channels(self) -> int:
        return sum(len(p.vols) for p in self.parts)

    @property
    def spatial_shape(self) -> tuple:
        return self.parts[0].shape

    @property
    def shape(self) -> tuple:
        return (self.n_channels, *self.spatial_shape)

    @property
    def device(self):
        return self.parts[0].vols[0].device

    def clone(self) -> "Pending":
        return Pending(parts=[p.copy() for p in self.parts], meta={k: (list(v) if isinstance(v, list) else v) for k, v in self.meta.items()})

    @staticmethod
    def cat(entries: Sequence["Pending"]) -> "Pending":
        return Pending(parts=[p for e in entries for p in e.parts])

    def tensor(self) -> torch.Tensor:
        """Materialise this entry alone (one launch per pass)."""
        from . import engine

        out = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        engine.execute(replay_parts(self.parts), [out[c] for c in range(self.n_channels)])
        return out


def as_pending(x) -> Pending:
    if isinstance(x, Pending):
        return x
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"cannot record transforms on {type(x)}")
    return Pending(x)


def _shape_of(x) -> tuple:
    return tuple(x.shape)


# --------------------------------------------------------------------------- protocol
class Transform:
    def __call__(self, data):
        raise NotImplementedError


class Randomizable:
    """MONAI ``Randomizable``: own ``np.random.RandomState`` in ``self.R`` †."""

    R: np.random.RandomState = np.random.RandomState()

    def set_random_state(self, seed: int | None = None, state: np.random.RandomState | None = None):
        if seed is not None:
            self.R = np.random.RandomState(int(seed) % MAX_SEED)
            return self
        if state is not None:
            if not isinstance(state, np.random.RandomState):
                raise TypeError(f"state must be None or a np.random.RandomState but is {type(state).__name__}.")
            self.R = state
            return self
        self.R = np.random.RandomState()
        return self

    def randomize(self, data=None) -> None:
        raise NotImplementedError


class MapTransform(Transform):
    def __init__(self, keys, allow_missing_keys: bool = False):
        self.keys = tuple(keys) if isinstance(keys, (list, tuple)) else (keys,)
        if not self.keys:
            raise ValueError("keys must be non empty.")
        self.allow_missing_keys = allow_missing_keys

    def key_iterator(self, data: dict):
        for k in self.keys:
            if k in data:
                yield k
            elif not self.allow_missing_keys:
                raise KeyError(f"Key `{k}` of transform `{type(self).__name__}` was missing in the data")

    def first_key(self, data: dict):
        for k in self.key_iterator(data):
            return k
        return ()


class RandomizableTransform(Randomizable, Transform):
    def __init__(self, prob: float = 1.0, do_transform: bool = True):
        self._do_transform = do_transform
        self.prob = min(max(prob, 0.0), 1.0)

    def randomize(self, data=None) -> None:
        self._do_transform = self.R.rand() < self.prob


def _per_key(value, keys, name="mode"):
    if isinstance(value, (list, tuple)):
        if len(value) != len(keys):
            raise ValueError(f"{name} must have one entry per key")
        return list(value)
    return [value] * len(keys)


def apply_transform(t: Callable, data):
    """``monai.transforms.apply_transform``: a list input maps the transform over its items."""
    if isinstance(data, (list, tuple)):
        return [apply_transform(t, d) for d in data]
    return t(data)


class Compose(Randomizable, Transform):
    """``monai.transforms.Compose`` †: sequential application, list outputs fan out, and
    ``set_random_state(seed)`` hands every Randomizable child ``R.randint(2**32)`` in order."""

    def __init__(self, transforms=None):
        if transforms is None:
            transforms = []
        if not isinstance(transforms, (list, tuple)):
            transforms = [transforms]
        self.transforms = tuple(transforms)
        self.set_random_state()  # fresh entropy, like MONAI without set_determinism (no draw from the global stream)

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed=seed, state=state)
        for t in self.transforms:
            if isinstance(t, Randomizable):
                t.set_random_state(seed=int(self.R.randint(MAX_SEED, dtype="uint32")))
        return self

    def randomize(self, data=None) -> None:
        for t in self.transforms:
            if isinstance(t, Randomizable):
                try:
                    t.randomize(data)
                except TypeError:
                    pass

    def flatten(self):
        out = []
        for t in self.transforms:
            if type(t) is Compose:
                out.extend(t.flatten().transforms)
            else:
                out.append(t)
        return Compose(out)

    def __len__(self):
        return len(self.flatten().transforms) if self.transforms else 0

    def __call__(self, data):
        for t in self.transforms:
            data = apply_transform(t, data)
        return data


class OneOf(Compose):
    """``monai.transforms.OneOf`` †: ``R.multinomial(1, weights).argmax()`` picks the member."""

    def __init__(self, transforms=None, weights=None):
        super().__init__(transforms)
        n = len(self.transforms)
        if n == 0:
            self.weights = []
        elif weights is None:
            self.weights = [1.0 / n] * n
        else:
            w = np.asarray(weights, dtype=float)
            if len(w) != n or (w < 0).any() or w.sum() <= 0:
                raise ValueError("weights must be non-negative, one per transform, and not all zero")
            self.weights = list(w / w.sum())

    def __call__(self, data):
        if len(self.transforms) == 0:
            return data
        index = int(self.R.multinomial(1, self.weights).argmax())
        return apply_transform(self.transforms[index], data)


class SomeOf(Compose):
    """``monai.transforms.SomeOf`` †: ``R.randint(min, max+1)`` members chosen by ``R.choice``."""

    def __init__(self, transforms=None, num_transforms=None, replace: bool = False, weights=None):
        super().__init__(transforms)
        n = len(self.transforms)
        if num_transforms is None:
            self.min_num_transforms = self.max_num_transforms = n
        elif isinstance(num_transforms, (tuple, list)):
            self.min_num_transforms, self.max_num_transforms = (int(x) for x in num_transforms)
        else:
            self.min_num_transforms = self.max_num_transforms = int(num_transforms)
        if self.max_num_transforms > n and not replace:
            raise ValueError("num_transforms cannot exceed the number of transforms without replacement")
        self.replace = replace
        self.weights = None if weights is None else list(np.asarray(weights, float) / np.sum(weights))

    def __call__(self, data):
        if len(self.transforms) == 0:
            return data
        sample_size = self.R.randint(self.min_num_transforms, self.max_num_transforms + 1)
        order = self.R.choice(len(self.transforms), sample_size, replace=self.replace, p=self.weights).tolist()
        for i in order:
            data = apply_transform(self.transforms[i], data)
        return data


# --------------------------------------------------------------------------- bookkeeping transforms
class Identityd(MapTransform):
    def __call__(self, data):
        return dict(data)


class EnsureTyped(MapTransform):
    """Outputs of the fused path are always float32 device tensors: recorded as a no-op."""

    def __init__(self, keys, dtype=None, track_meta=None, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)

    def __call__(self, data):
        return dict(data)


class ToTensord(EnsureTyped):
    """``monai.transforms.ToTensord`` † (same conversion as :class:`EnsureTyped` here)."""



class SelectItemsd(MapTransform):
    def __call__(self, data):
        return {k: data[k] for k in self.key_iterator(data)}


class Lambdad(MapTransform):
    def __init__(self, keys, func, allow_missing_keys: bool = False):
        super().__init__(keys, allow_missing_keys)
        self.func = func

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            d[k] = self.func(d[k])
        return d


class ConcatItemsd(MapTransform):
    """Channel concatenation: the recorded chains are concatenated, voxels are written once."""

    def __init__(self, keys, name: str, dim: int = 0, allow_missing_keys: bool = False):
        super().__init__(keys, allow_missing_keys)
        if dim != 0:
            raise NotImplementedError("ConcatItemsd: only channel concatenation (dim=0) is on the fused path")
        self.name = name

    def __call__(self, data):
        d = dict(data)
        entries = [as_pending(d[k]) for k in self.key_iterator(d)]
        if not entries:
            return d
        shapes = {e.spatial_shape for e in entries}
        if len(shapes) != 1:
            raise ValueError(f"ConcatItemsd: spatial shapes differ: {shapes}")
        d[self.name] = Pending.cat(entries)
        return d


class ExposeTransformKeyMetad(Transform):
    """/root/reference/adell_mri/utils/monai_transforms/generic_data_ops.py:75-119: exposes a value
    recorded by an applied transform (here: ``extra_info.cropped`` of the last RandSpatialCrop)."""

    def __init__(self, key: str, transform_class: str, nested_pattern: Sequence[str], output_key: str | None = None):
        self.key, self.transform_class, self.nested_pattern = key, transform_class, list(nested_pattern)
        self.output_key = output_key if output_key is not None else "box_" + key   # the reference's default

    def __call__(self, data):
        d = dict(data)
        ops = d[self.key].meta.get("applied_operations", []) if isinstance(d[self.key], Pending) else []
        for op in ops:
            if op["class"] == self.transform_class:
                v = op
                for p in self.nested_pattern:
                    v = v[p]
                d[self.output_key] = v
        return d


# --------------------------------------------------------------------------- integer geometry
class SpatialPadd(MapTransform):
    """``monai.transforms.SpatialPadd`` (method symmetric, constant 0) †."""

    def __init__(self, keys, spatial_size, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        self.spatial_size = [int(x) for x in spatial_size]

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.spatial_pad(self.spatial_size)
        return d


class CenterSpatialCropd(MapTransform):
    def __init__(self, keys, roi_size, allow_missing_keys: bool = False):
        super().__init__(keys, allow_missing_keys)
        self.roi_size = [int(x) for x in roi_size]

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.center_crop(self.roi_size)
        return d


class SpatialCropd(MapTransform):
    def __init__(self, keys, roi_start, roi_size, allow_missing_keys: bool = False):
        super().__init__(keys, allow_missing_keys)
        self.roi_start, self.roi_size = [int(x) for x in roi_start], [int(x) for x in roi_size]

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.crop(self.roi_start, self.roi_size)
        return d


class RandSpatialCropd(Randomizable, MapTransform):
    """``monai.transforms.RandSpatialCropd`` †: randomised once on the first key's shape, the same
    window for every key; records ``extra_info.cropped = [start0, size0-end0, start1, ...]``."""

    def __init__(self, keys, roi_size, max_roi_size=None, random_center: bool = True, random_size: bool = False,
                 allow_missing_keys: bool = False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        self.roi_size = list(roi_size) if isinstance(roi_size, (list, tuple)) else [roi_size] * 3
        self.max_roi_size = max_roi_size
        self.random_center, self.random_size = random_center, random_size
        self._size, self._start = None, None

    def randomize(self, img_size) -> None:
        size = [d if (r is None or r <= 0) else min(int(r), d) for d, r in zip(img_size, self.roi_size)]
        if self.random_size:
            mx = list(img_size) if self.max_roi_size is None else [min(int(m), d) for m, d in zip(self.max_roi_size, img_size)]
            if any(a > b for a, b in zip(size, mx)):
                raise ValueError(f"min ROI size: {size} is larger than max ROI size: {mx}.")
            size = [int(self.R.randint(low=size[i], high=mx[i] + 1)) for i in range(len(img_size))]
        self._size = size
        if self.random_center:
            self._start = [int(self.R.randint(low=0, high=d - s + 1)) if d > s else 0 for d, s in zip(img_size, size)]
        else:
            self._start = [max(d // 2 - s // 2, 0) for d, s in zip(img_size, size)]

    def __call__(self, data):
        d = dict(data)
        first = self.first_key(d)
        if first == ():
            return d
        d[first] = as_pending(d[first])
        self.randomize(d[first].spatial_shape)
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            shape = d[k].spatial_shape
            d[k].plan.crop(self._start, self._size)
            cropped = []
            for dim, s, z in zip(shape, self._start, self._size):
                cropped += [int(s), int(dim - min(s + z, dim))]
            d[k].meta.setdefault("applied_operations", []).append(
                {"class": "RandSpatialCrop", "extra_info": {"cropped": cropped}})
        return d


class FgBgToIndicesd(MapTransform):
    """``monai.transforms.FgBgToIndicesd`` †: flat indices of foreground (>0) and background
    voxels of a (materialised, cached) label volume.  For a label on the device the two lists STAY there
    (int64 tensors: their lengths are host integers); for a host label they are numpy arrays."""

    def __init__(self, keys, fg_postfix: str = "_fg_indices", bg_postfix: str = "_bg_indices", allow_missing_keys: bool = False,
                 on_device: bool | None = None):
        super().__init__(keys, allow_missing_keys)
        self.fg_postfix, self.bg_postfix = fg_postfix, bg_postfix
        self.on_device = on_device   # None: where the label lives

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            lab = d[k].tensor() if isinstance(d[k], Pending) else torch.as_tensor(d[k])
            flat = (lab > 0).any(dim=0).reshape(-1)
            fg, bg = torch.nonzero(flat).reshape(-1), torch.nonzero(~flat).reshape(-1)
            if not (self.on_device if self.on_device is not None else lab.is_cuda):
                fg, bg = fg.cpu().numpy(), bg.cpu().numpy()
            # (device lists: only their lengths are known to the host, which is all the crop draws need —
            # RandCropByPosNegLabeld then selects the centres with adell_posneg_starts)
            d[k + self.fg_postfix], d[k + self.bg_postfix] = fg, bg
        return d


def _posneg_starts(picks, label_shape, size, device) -> torch.Tensor:
    """``[len(picks), 3]`` int32 crop starts on ``device`` for host draws ``picks = [(index list tensor, entry)]``."""
    import ctypes as C

    from . import _lib, engine

    n = len(picks)
    arr = np.zeros(n, np.dtype(_lib.PosNeg))
    arr["indices"] = [lst.data_ptr() for lst, _ in picks]
    arr["pick"] = [p for _, p in picks]
    arr["shape"] = [int(x) for x in label_shape]
    arr["size"] = [int(x) for x in size]
    out = torch.empty((n, 3), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        desc = engine._stage(arr.view(np.uint8).reshape(-1), device)
        _lib.check(_lib.load().adell_posneg_starts(desc.data_ptr(), n, out.data_ptr(),
                                                   C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "adell_posneg_starts")
    out._adell_keep = [lst for lst, _ in picks]   # the lists must outlive the select kernel's launch
    return out


class RandCropByPosNegLabeld(Randomizable, MapTransform):
    """``monai.transforms.RandCropByPosNegLabeld`` († generate_pos_neg_label_crop_centers +
    correct_crop_centers(allow_smaller=True)): returns a LIST of ``num_samples`` dicts."""

    def __init__(self, keys, label_key, spatial_size, pos: float = 1.0, neg: float = 1.0, num_samples: int = 1,
                 fg_indices_key=None, bg_indices_key=None, allow_smaller: bool = False, allow_missing_keys: bool = False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        if pos < 0 or neg < 0 or pos + neg == 0:
            raise ValueError("pos and neg must be non-negative and not both zero")
        self.label_key, self.spatial_size = label_key, [int(x) for x in spatial_size]
        self.pos_ratio = pos / (pos + neg)
        self.num_samples = num_samples
        self.fg_indices_key, self.bg_indices_key = fg_indices_key, bg_indices_key
        self.allow_smaller = allow_smaller
        self.centers = None

    def randomize(self, label_shape, fg_indices, bg_indices) -> None:
        size = [min(int(s), int(d)) if s > 0 else int(d) for s, d in zip(self.spatial_size, label_shape)]
        if not isinstance(fg_indices, torch.Tensor):
            fg_indices, bg_indices = np.asarray(fg_indices), np.asarray(bg_indices)
        pos_ratio = self.pos_ratio
        if len(fg_indices) == 0 or len(bg_indices) == 0:
            if len(fg_indices) == 0 and len(bg_indices) == 0:
                raise ValueError("No sampling location available.")
            pos_ratio = 0 if len(fg_indices) == 0 else 1
        centers = []
        if isinstance(fg_indices, torch.Tensor):
            # device lists: the same two draws per crop; the entry itself is looked up on the device
            self._picks = []
            for _ in range(self.num_samples):
                use_fg = self.R.rand() < pos_ratio
                lst = fg_indices if use_fg else bg_indices
                self._picks.append((lst, int(self.R.randint(len(lst)))))
            self.centers, self._size = None, size
            return
        for _ in range(self.num_samples):
            indices_to_use = fg_indices if self.R.rand() < pos_ratio else bg_indices
            idx = indices_to_use[self.R.randint(len(indices_to_use))]
            center = np.unravel_index(idx, label_shape)
            valid_start = np.floor_divide(size, 2)
            valid_end = np.subtract(np.array(label_shape) + 1, np.array(size) / 2).astype(np.uint16)
            for i, vs in enumerate(valid_start):
                if vs == valid_end[i]:
                    valid_end[i] += 1
            centers.append([int(min(max(c, vs), ve - 1)) for c, vs, ve in zip(center, valid_start, valid_end)])
        self.centers, self._size = centers, size

    def __call__(self, data):
        d = dict(data)
        lab = as_pending(d[self.label_key])
        fg = d.get(self.fg_indices_key) if self.fg_indices_key else None
        bg = d.get(self.bg_indices_key) if self.bg_indices_key else None
        if fg is None or bg is None:
            t = lab.tensor()
            flat = (t > 0).any(dim=0).reshape(-1)
            fg, bg = torch.nonzero(flat).reshape(-1).cpu().numpy(), torch.nonzero(~flat).reshape(-1).cpu().numpy()
        self.randomize(lab.spatial_shape, fg, bg)
        out = []
        if self.centers is None:
            # centres selected on the device: adell_posneg_starts writes the crop starts where K1 reads them
            # (ADELL_F_WIN_DEV); nothing about the label's content ever reaches the host
            win = _posneg_starts(self._picks, lab.spatial_shape, self._size, lab.device)
            for i in range(self.num_samples):
                r = dict(d)
                for k in self.key_iterator(d):
                    e = as_pending(d[k]).clone()
                    e.plan.crop_from_device(win[i], self._size)
                    r[k] = e
                out.append(r)
            return out
        for center in self.centers:
            r = dict(d)
            for k in self.key_iterator(d):
                e = as_pending(d[k]).clone()
                starts = [max(int(c) - int(s) // 2, 0) for c, s in zip(center, self._size)]
                e.plan.crop(starts, self._size)
                r[k] = e
            out.append(r)
        return out


class RandFlipd(RandomizableTransform, MapTransform):
    """``monai.transforms.RandFlipd`` †: one ``R.rand() < prob`` per call; ``torch.flip``."""

    def __init__(self, keys, prob: float = 0.1, spatial_axis=None, allow_missing_keys: bool = False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        if spatial_axis is None:
            spatial_axis = (0, 1, 2)
        self.spatial_axis = [spatial_axis] if isinstance(spatial_axis, (int, np.integer)) else list(spatial_axis)

    def __call__(self, data):
        d = dict(data)
        self.randomize(None)
        # (an idle call records the empty flip: every sample of a batch then carries the same chain, and
        # safe_collate composes the whole batch as one group)
        mask = tuple(self._do_transform and a in self.spatial_axis for a in range(3))
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.flip(mask)
        return d


# --------------------------------------------------------------------------- resampling
class RandAffined(RandomizableTransform, MapTransform):
    """``monai.transforms.RandAffined`` †: same constructor vocabulary as the reference uses
    (rotate / shear / translate / scale ranges, per-key ``mode``, ``padding_mode`` — MONAI's
    default "reflection" when the reference leaves it unset), three identically seeded streams,
    one shared grid for every key (see :class:`~adell_mri_b200.sampling.RandAffineSampler`)."""

    def __init__(self, keys, spatial_size=None, prob: float = 0.1, rotate_range=None, shear_range=None,
                 translate_range=None, scale_range=None, mode="bilinear", padding_mode="reflection",
                 cache_grid: bool = False, device=None, allow_missing_keys: bool = False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        if spatial_size is not None:
            raise NotImplementedError("RandAffined(spatial_size=...) is not used by the reference and not on the fused path")
        self.sampler = RandAffineSampler(prob, rotate_range, shear_range, translate_range, scale_range)
        self.mode = _per_key(mode, self.keys, "mode")
        self.padding_mode = _per_key(padding_mode, self.keys, "padding_mode")
        self.last_affine = None

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        self.sampler.set_random_state(seed, state)
        self.R = self.sampler.R
        return self

    def __call__(self, data):
        d = dict(data)
        keys = list(self.key_iterator(d))
        fired, p = self.sampler.draw(n_keys=len(keys))
        self._do_transform = bool(fired)
        A = geometry.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"])[0] if fired else _EYE4
        self.last_affine = A if fired else None
        for k, mode, pad in zip(self.keys, self.mode, self.padding_mode):
            if k not in d:
                continue
            d[k] = as_pending(d[k])
            d[k].plan.affine(A, mode, pad, fired=bool(fired))   # idle calls record "not fired" (same chain for the whole batch)
        return d


# --------------------------------------------------------------------------- intensity
class _RandIntensityd(RandomizableTransform, MapTransform):
    """Shared draw order of the Rand*Intensityd family †: the dict transform's own ``R.rand()``
    gate, then the wrapped array transform's ``R.rand()`` (prob 1.0) and its parameter draw from
    an identically seeded second stream."""

    def __init__(self, keys, prob, allow_missing_keys=False):
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        self.R_inner = np.random.RandomState()

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        if state is not None:
            self.R_inner = state
        else:
            self.R_inner = np.random.RandomState(seed)
        return self


class RandScaleIntensityd(_RandIntensityd):
    """``v * (1 + factor)``, ``factor ~ U(-factors, factors)`` †."""

    def __init__(self, keys, factors, prob: float = 0.1, allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        self.factors = (min(-factors, factors), max(-factors, factors)) if not isinstance(factors, (tuple, list)) else (min(factors), max(factors))
        self.factor = None

    def __call__(self, data):
        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        self.R_inner.rand()
        self.factor = self.R_inner.uniform(low=self.factors[0], high=self.factors[1])
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.intensity(scale=float(np.float32(1 + self.factor)))
        return d


class RandShiftIntensityd(_RandIntensityd):
    """``v + offset``, ``offset ~ U(-offsets, offsets)`` †."""

    def __init__(self, keys, offsets, prob: float = 0.1, allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        self.offsets = (min(-offsets, offsets), max(-offsets, offsets)) if not isinstance(offsets, (tuple, list)) else (min(offsets), max(offsets))
        self.offset = None

    def __call__(self, data):
        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        self.R_inner.rand()
        self.offset = self.R_inner.uniform(low=self.offsets[0], high=self.offsets[1])
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            d[k].plan.intensity(offset=float(np.float32(self.offset)))
        return d


class RandStdShiftIntensityd(_RandIntensityd):
    """``monai.transforms.RandStdShiftIntensityd`` †: ``v + factor * std(v)``, ``factor ~ U(-factors,
    factors)``, the same factor for every key, ``std`` = population standard deviation of each key's
    whole array (``torch.std(unbiased=False)``), computed by the device statistics kernel (fp64
    accumulation) and applied through the fused ``{scale, offset}`` pre map.  The input is
    materialised first (the statistic is taken over the current volume)
    (/root/reference/adell_mri/transform_factory/augmentations.py:66-76,219-232)."""

    def __init__(self, keys, factors, prob: float = 0.1, nonzero: bool = False, channel_wise: bool = False,
                 allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        if nonzero or channel_wise:
            raise NotImplementedError("RandStdShiftIntensityd: nonzero / channel_wise are not used by the reference")
        self.factors = (min(-factors, factors), max(-factors, factors)) if not isinstance(factors, (tuple, list)) else (min(factors), max(factors))
        self.factor = None

    def __call__(self, data):
        from . import stats

        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        self.R_inner.rand()
        self.factor = self.R_inner.uniform(low=self.factors[0], high=self.factors[1])
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            x = x.contiguous()
            ms = stats.meanstd([x.reshape(-1)], raw_std=True)            # [1, 2] on the device
            pre = torch.ones(x.shape[0], 2, dtype=torch.float32, device=x.device)
            pre[:, 1] = ms[0, 1] * self.factor                           # fp32 tensor * python scalar, like MONAI
            p = as_pending(x.to(torch.float32) if x.dtype != torch.float32 else x)
            p.plan.intensity_from_device(pre)
            d[k] = p
        return d


class RandAdjustContrastd(_RandIntensityd):
    """``monai.transforms.RandAdjustContrastd`` †: ``((v - min) / (range + 1e-7)) ** gamma * range +
    min`` with ``gamma ~ U(gamma)`` shared by every key and ``min`` / ``range`` of each key's whole
    array.  The power law is not linear, so it cannot ride in the gather's intensity map: it runs
    as one pointwise device pass (``adell_minmax`` + ``adell_gamma_map``) over the materialised
    input (/root/reference/adell_mri/transform_factory/augmentations.py:66-76,219-232)."""

    def __init__(self, keys, prob: float = 0.1, gamma=(0.5, 4.5), invert_image: bool = False, retain_stats: bool = False,
                 allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        if invert_image or retain_stats:
            raise NotImplementedError("RandAdjustContrastd: invert_image / retain_stats are not used by the reference")
        if isinstance(gamma, (int, float)):
            if gamma <= 0.5:
                raise ValueError("if gamma is a number, must greater than 0.5 and value is picked from (0.5, gamma)")
            self.gamma = (0.5, gamma)
        else:
            self.gamma = (min(gamma), max(gamma))
        self.gamma_value = None

    def __call__(self, data):
        from . import stats

        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        self.R_inner.rand()
        self.gamma_value = self.R_inner.uniform(low=self.gamma[0], high=self.gamma[1])
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            flat = [x.contiguous().reshape(-1)]
            mm = stats.minmax(flat)
            d[k] = stats.gamma_map(flat, mm, self.gamma_value)[0].reshape(x.shape)
        return d


class Resized(MapTransform):
    """``monai.transforms.Resized`` † (``size_mode="all"``, no anti-aliasing): ``F.interpolate`` with
    MONAI's default ``mode="area"`` (adaptive average pooling) or ``"nearest"``, per key; a no-op when
    the shape already matches.  Runs on the device (``adell_resize``, bit-identical to ATen's CPU
    kernel) over the materialised input and hands a fresh entry on
    (/root/reference/adell_mri/transform_factory/augmentations.py:427-444: the SSL scaled crop;
    /root/reference/adell_mri/transform_factory/transforms.py:157-167,455-462: ``resize_size``)."""

    def __init__(self, keys, spatial_size, size_mode: str = "all", mode="area", align_corners=None,
                 anti_aliasing: bool = False, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        if size_mode != "all" or anti_aliasing:
            raise NotImplementedError("Resized: only size_mode='all' without anti-aliasing is used by the reference")
        self.spatial_size = [int(x) for x in spatial_size]
        n = len(self.keys)
        self.mode = [mode] * n if isinstance(mode, str) else list(mode)
        for m in self.mode:
            if m not in ("area", "nearest"):
                raise NotImplementedError(f"Resized mode '{m}' is not on the device path (area / nearest are)")

    def __call__(self, data):
        from . import stats

        d = dict(data)
        for k, m in zip(self.keys, self.mode):
            if k not in d:
                if self.allow_missing_keys:
                    continue
                raise KeyError(k)
            x = d[k]
            shape = tuple(x.shape[1:])
            if shape == tuple(self.spatial_size):
                continue
            t = (x.tensor() if isinstance(x, Pending) else x).to(torch.float32)
            outs = stats.resize([t[c] for c in range(t.shape[0])], self.spatial_size, m)
            d[k] = torch.stack(outs, 0) if len(outs) > 1 else outs[0][None]
        return d


def _axis_spacing(entry, d: dict, key: str) -> np.ndarray:
    """Voxel spacing of ``d[key]`` along its three axes: from ``d[f"{key}_spacing"]`` (3 numbers), else from an
    ``affine`` (attribute of the tensor, like MONAI's MetaTensor, or ``d[f"{key}_meta_dict"]["affine"]``), which
    must be axis-aligned (what ``Orientationd("RAS")`` leaves for a scanner-aligned acquisition)."""
    if f"{key}_spacing" in d:
        return np.asarray(d[f"{key}_spacing"], np.float64).reshape(3)
    aff = getattr(entry, "affine", None)
    if aff is None and isinstance(getattr(entry, "meta", None), dict):
        aff = entry.meta.get("affine")
    if aff is None and isinstance(d.get(f"{key}_meta_dict"), dict):
        aff = d[f"{key}_meta_dict"].get("affine")
    if aff is None:
        raise KeyError(f"Spacingd: no spacing for '{key}' (give '{key}_spacing', an .affine attribute or '{key}_meta_dict')")
    aff = np.asarray(aff, np.float64)[:3, :3]
    if np.abs(aff - np.diag(np.diag(aff))).max() > 1e-6 * np.abs(aff).max() or (np.diag(aff) <= 0).any():
        raise NotImplementedError("Spacingd on the fused path needs an axis-aligned affine with positive spacings (oblique "
                                  "acquisitions need the general SpatialResample of the loading stage)")
    return np.diag(aff).copy()


class Spacingd(MapTransform):
    """``monai.transforms.Spacingd(keys, pixdim, mode)`` † for axis-aligned volumes
    (/root/reference/adell_mri/transform_factory/transforms.py:133-140,444-454: the cached stage's resample to
    ``target_spacing``).  MONAI's ``Spacing`` builds the output affine with the new pixel size and the SAME
    origin (voxel 0 stays in place), sizes the output with ``compute_shape_offset``
    (``round((S - 1) * spacing / pixdim + 1)`` per axis), and resamples index-to-index
    (``align_corners=False``, ``padding_mode="border"``): output voxel ``o`` reads source index
    ``o * pixdim / spacing``.  Recorded as one K1 resample item (:meth:`BatchPlan.resample_to`).
    † restated from MONAI 1.3-1.6 (``Spacing.__call__`` -> ``SpatialResample``); MONAI evaluates the grid in
    float64 by default, K1 in fp32: trilinear results agree within the path's 1e-4, nearest ones except at
    coordinates within ~1e-5 voxel of a rounding tie (exactly representable steps such as 0.5 / 2 are exact)."""

    def __init__(self, keys, pixdim, diagonal: bool = False, mode="bilinear", padding_mode="border",
                 align_corners: bool = False, dtype=None, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        if diagonal or align_corners:
            raise NotImplementedError("Spacingd: diagonal=True / align_corners=True are not used by the reference")
        self.pixdim = np.asarray(pixdim, np.float64).reshape(-1)
        self.mode = _per_key(mode, self.keys, "mode")
        self.padding_mode = _per_key(padding_mode, self.keys, "padding_mode")

    @staticmethod
    def output_size(shape, spacing, pixdim):
        """compute_shape_offset for an axis-aligned affine: ``np.round(ptp + 1)`` per axis (round half to even)."""
        ext = (np.asarray(shape, np.float64) - 1.0) * (np.asarray(spacing, np.float64) / np.asarray(pixdim, np.float64))
        return np.round(ext + 1.0).astype(np.int64)

    def __call__(self, data):
        d = dict(data)
        for k, mode, pad in zip(self.keys, self.mode, self.padding_mode):
            if k not in d:
                if self.allow_missing_keys:
                    continue
                raise KeyError(k)
            spacing = _axis_spacing(d[k], d, k)
            pixdim = np.where(self.pixdim[:3] > 0, self.pixdim[:3], spacing) if self.pixdim.size >= 3 else np.full(3, self.pixdim[0])
            d[k] = as_pending(d[k])
            shape = d[k].spatial_shape
            out = self.output_size(shape, spacing, pixdim)
            if np.allclose(pixdim, spacing, rtol=0, atol=1e-3) and tuple(out) == tuple(shape):
                continue   # MONAI: affine unchanged -> the image is handed on as it is
            d[k].plan.resample_to(pixdim / spacing, out, mode, pad)
            d[f"{k}_spacing"] = pixdim.copy()
        return d


class RandRicianNoised(_RandIntensityd):
    """``monai.transforms.RandRicianNoised`` †: the dict transform's ``R.rand()`` gate, then PER KEY the
    wrapped ``RandRicianNoise(prob=1.0)``: its own ``R.rand()``, ``sigma ~ U(0, std)`` (``sample_std``)
    and two ``R.normal(mean, sigma, size=shape)`` volumes (float64 -> float32) from the identically
    seeded second stream; ``out = sqrt((v + n1)**2 + n2**2)``.  Not linear, so it cannot ride in the
    gather's intensity map: one streaming device pass (``adell_rician_map``) over the materialised
    input (/root/reference/adell_mri/modules/augmentations.py:53,86,117 ``rician_noise``;
    /root/reference/adell_mri/transform_factory/augmentations.py:81-91)."""

    def __init__(self, keys, prob: float = 0.1, mean: float = 0.0, std: float = 1.0, channel_wise: bool = False,
                 relative: bool = False, sample_std: bool = True, allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        if channel_wise or relative:
            raise NotImplementedError("RandRicianNoised: channel_wise / relative are not used by the reference")
        self.mean, self.std, self.sample_std = mean, std, sample_std
        self.noise = {}

    def __call__(self, data):
        from . import stats

        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        self.noise = {}
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            self.R_inner.rand()
            std = self.R_inner.uniform(0, self.std) if self.sample_std else self.std
            n1 = self.R_inner.normal(self.mean, std, size=tuple(x.shape)).astype(np.float32)
            n2 = self.R_inner.normal(self.mean, std, size=tuple(x.shape)).astype(np.float32)
            self.noise[k] = (n1, n2)
            up = [torch.from_numpy(n).pin_memory().to(x.device, non_blocking=True) for n in (n1, n2)]
            d[k] = stats.rician_map(x.to(torch.float32), up[0], up[1])
        return d


class RandGaussianNoised(_RandIntensityd):
    """``monai.transforms.RandGaussianNoised`` †: sigma ~ U(0, std) (``sample_std``), ONE noise
    volume of the first key's shape drawn from ``R.normal`` on the host (float64 -> float32) and
    added to every key.  In ``noise="philox"`` mode only sigma is drawn on the host and the
    normal deviates come from the device-side Philox generator (not stream-comparable)."""

    def __init__(self, keys, prob: float = 0.1, mean: float = 0.0, std: float = 0.1, sample_std: bool = True,
                 allow_missing_keys: bool = False, **_):
        super().__init__(keys, prob, allow_missing_keys)
        self.mean, self.std, self.sample_std = mean, std, sample_std
        self.noise = None
        self._philox_calls = 0

    def __call__(self, data):
        d = dict(data)
        self.randomize(None)
        if not self._do_transform:
            return d
        first = self.first_key(d)
        if first == ():
            return d
        d[first] = as_pending(d[first])
        shape = d[first].shape
        self.R_inner.rand()
        std = self.R_inner.uniform(0, self.std) if self.sample_std else self.std
        if _MODE["noise"] == "philox":
            if self.mean != 0.0:
                raise NotImplementedError("philox noise supports mean=0 only")
            self._philox_calls += 1
            seed = int(self.R_inner.randint(MAX_SEED, dtype="uint32"))
            offs = np.arange(shape[0], dtype=np.uint64) * np.uint64(int(np.prod(shape[1:])))
            for k in self.key_iterator(d):
                d[k] = as_pending(d[k])
                d[k].plan.add_philox_noise(np.float32(std), seed=seed, offset=offs)
            return d
        noise = self.R_inner.normal(self.mean, std, size=shape).astype(np.float32)
        self.noise = noise
        for k in self.key_iterator(d):
            d[k] = as_pending(d[k])
            if d[k].shape != shape:
                raise ValueError("RandGaussianNoised: all keys must share the first key's shape")
            dev = d[k].device
            nz = torch.from_numpy(noise)
            nz = nz.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else nz
            d[k].plan.add_noise([nz[c] for c in range(shape[0])])
        return d


class ScaleIntensityd(MapTransform):
    """``monai.transforms.ScaleIntensityd`` († min-max to [minv, maxv], or ``v*(1+factor)``) on the
    device statistics kernels, bit-exact fp32 operation order.  A cached-stage transform: it
    materialises its input (the reference runs it before the CacheDataset cut as well)."""

    def __init__(self, keys, minv=0.0, maxv=1.0, factor=None, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        self.minv, self.maxv, self.factor = minv, maxv, factor

    def __call__(self, data):
        from . import _lib, stats

        d = dict(data)
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            vols = [x[c].contiguous() for c in range(x.shape[0])]
            flat = [x.contiguous().reshape(-1)]  # MONAI scales over the whole array (channel_wise=False)
            if self.minv is not None or self.maxv is not None:
                mm = stats.minmax(flat)
                coefs = stats.scaler_coefs(mm, _lib.SCALER_MINMAX, self.minv, self.maxv)
                out = stats.intensity_map(flat, coefs)[0].reshape(x.shape)
            elif self.factor is not None:
                out = (x.to(torch.float32) * np.float32(1 + self.factor)).to(torch.float32)
            else:
                raise ValueError("Incompatible values: minv=None or maxv=None and factor=None.")
            d[k] = out
            del vols
        return d


class NormalizeIntensityd(MapTransform):
    """``monai.transforms.NormalizeIntensityd`` († ``(x - mean) / std``, population std, a zero std
    replaced by 1; ``nonzero`` restricts the statistics — and the update — to voxels != 0) on the
    device statistics kernels (fp64 accumulation: within 1e-6 of torch's fp32 reduction).  Named by
    the north star; the reference's own pipelines use min-max scaling."""

    def __init__(self, keys, subtrahend=None, divisor=None, nonzero: bool = False, channel_wise: bool = False,
                 allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        if subtrahend is not None or divisor is not None:
            raise NotImplementedError("NormalizeIntensityd: explicit subtrahend / divisor are not on the device path")
        self.nonzero, self.channel_wise = nonzero, channel_wise

    def __call__(self, data):
        from . import _lib, stats

        d = dict(data)
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            x = x.contiguous()
            vols = [x[c].reshape(-1) for c in range(x.shape[0])] if self.channel_wise else [x.reshape(-1)]
            ms = stats.meanstd(vols, nonzero=self.nonzero)
            out = torch.stack([o for o in stats.intensity_map(vols, stats.scaler_coefs(ms, _lib.SCALER_ZSCORE, 0.0, 0.0))]).reshape(x.shape)
            if self.nonzero:
                out = torch.where(x != 0, out, x.to(torch.float32))
            d[k] = out
        return d


class ScaleIntensityRangePercentilesd(MapTransform):
    """``monai.transforms.ScaleIntensityRangePercentilesd`` †: ``a_min/a_max = np.percentile(x, q)``
    ('linear' method, exact via the radix-select kernels), then ScaleIntensityRange."""

    def __init__(self, keys, lower: float, upper: float, b_min, b_max, clip: bool = False, relative: bool = False,
                 channel_wise: bool = False, allow_missing_keys: bool = False, **_):
        super().__init__(keys, allow_missing_keys)
        if relative or b_min is None or b_max is None:
            raise NotImplementedError("ScaleIntensityRangePercentilesd: relative / open output ranges are not on the device path")
        if not (0 <= lower <= 100 and 0 <= upper <= 100):
            raise ValueError("Percentiles must be in the range [0, 100]")
        self.lower, self.upper, self.b_min, self.b_max, self.clip, self.channel_wise = lower, upper, b_min, b_max, clip, channel_wise

    def __call__(self, data):
        from . import _lib, stats

        d = dict(data)
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            x = x.contiguous()
            vols = [x[c].reshape(-1) for c in range(x.shape[0])] if self.channel_wise else [x.reshape(-1)]
            pct = stats.percentiles(vols, [self.lower, self.upper])
            coefs = stats.scaler_coefs(pct, _lib.SCALER_RANGE, self.b_min, self.b_max)
            outs = stats.intensity_map(vols, coefs, clip=(self.b_min, self.b_max) if self.clip else None)
            d[k] = torch.stack(outs).reshape(x.shape)
        return d


def not_on_fused_path(name: str):
    def ctor(*args, **kwargs):
        raise NotImplementedError(
            f"{name} is outside the fused GPU hot path (stencil / FFT / polynomial field / non-affine warp; "
            "see DESIGN.md, out of scope)")
    return ctor


RandGibbsNoised = not_on_fused_path("RandGibbsNoised")
RandBiasFieldd = not_on_fused_path("RandBiasFieldd")
RandGaussianSmoothd = not_on_fused_path("RandGaussianSmoothd")
RandGridDistortiond = not_on_fused_path("RandGridDistortiond")
RandSimulateLowResolutiond = not_on_fused_path("RandSimulateLowResolutiond")


def all_flip_combinations(flip_axis):
    out = []
    for i in range(len(flip_axis)):
        out.extend(itertools.combinations(flip_axis, i + 1))
    return out
