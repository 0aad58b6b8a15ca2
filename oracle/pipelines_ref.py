"""CPU ORACLE — eager restatement of the reference's augmentation pipelines (test infrastructure
only; never imported by the product package).

The reference builds ``monai.transforms.Compose`` pipelines in
/root/reference/adell_mri/transform_factory/augmentations.py:19-178 (unet), :181-320 (class),
:391-516 (ssl) and /root/reference/adell_mri/modules/augmentations.py:165-256 (workhorse), and
MONAI executes them eagerly, one materialised tensor per op.  MONAI (pinned 1.6.0) is not
importable here, so the few MONAI classes those builders use are restated below († = from MONAI
1.3-1.6 upstream behaviour) on top of oracle/monai_restated.py, as plain eager torch-CPU code:
every op produces a new tensor exactly as the reference would.

PARITY STATUS: unpinned against MONAI itself; the product's lazy transforms are
compared against these eager ones on identical seeds.
"""

from __future__ import annotations

import itertools

import numpy as np
import torch

from oracle import monai_restated as M

MAX_SEED = np.iinfo(np.uint32).max + 1


class Rnd:
    """Randomizable †."""

    def __init__(self):
        self.R = np.random.RandomState()

    def seed(self, s):
        self.R = np.random.RandomState(s)
        return self


def _seed_children(R, children):
    for c in children:
        if hasattr(c, "seed"):
            c.seed(int(R.randint(MAX_SEED, dtype="uint32")))


def run(t, data):
    if isinstance(data, list):
        return [run(t, d) for d in data]
    return t(data)


class Chain(Rnd):
    """Compose †."""

    def __init__(self, ts):
        super().__init__()
        self.ts = list(ts)

    def seed(self, s):
        super().seed(s)
        _seed_children(self.R, self.ts)
        return self

    def __call__(self, d):
        for t in self.ts:
            d = run(t, d)
        return d


class PickOne(Chain):
    """OneOf †: multinomial(1, uniform weights).argmax()."""

    def __call__(self, d):
        w = [1.0 / len(self.ts)] * len(self.ts)
        return run(self.ts[int(self.R.multinomial(1, w).argmax())], d)


class PickSome(Chain):
    """SomeOf(num_transforms=n) †."""

    def __init__(self, ts, n):
        super().__init__(ts)
        self.n = n

    def __call__(self, d):
        k = self.R.randint(self.n, self.n + 1)
        for i in self.R.choice(len(self.ts), k, replace=False, p=None).tolist():
            d = run(self.ts[i], d)
        return d


class Same:
    def __call__(self, d):
        return dict(d)


class AffineD(Rnd):
    """RandAffined † (three identically seeded streams, one grid for all keys)."""

    def __init__(self, keys, prob, modes, padding="reflection", **ranges):
        super().__init__()
        self.keys, self.modes, self.padding = list(keys), list(modes), padding
        self.draws = M.RandAffinedDraws(prob, n_keys=len(self.keys), **ranges)

    def seed(self, s):
        super().seed(s)
        self.draws.set_random_state(s)
        return self

    def __call__(self, d):
        d = dict(d)
        fired, p = self.draws.draw()
        if fired:
            A = M.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"])
            for k, m in zip(self.keys, self.modes):
                d[k] = M.affine_resample(d[k].to(torch.float32), A, m, self.padding)
        else:
            for k in self.keys:
                d[k] = d[k].to(torch.float32)
        return d


class FlipD(Rnd):
    def __init__(self, keys, prob, axes):
        super().__init__()
        self.keys, self.prob, self.axes = list(keys), prob, list(axes)

    def __call__(self, d):
        d = dict(d)
        if self.R.rand() < self.prob:
            for k in self.keys:
                d[k] = M.flip(d[k], self.axes)
        return d


class RandCropD(Rnd):
    """RandSpatialCropd † + the ``extra_info.cropped`` record the reference exposes as box_1/2."""

    def __init__(self, keys, roi, random_size=False):
        super().__init__()
        self.keys, self.roi, self.random_size = list(keys), list(roi), random_size

    def __call__(self, d):
        d = dict(d)
        shape = tuple(d[self.keys[0]].shape[1:])
        starts, sizes = M.rand_spatial_crop_draw(self.R, shape, self.roi, self.random_size)
        for k in self.keys:
            full = tuple(d[k].shape[1:])
            d[k] = M.crop(d[k], starts, sizes)
            d.setdefault("_cropped", {})[k] = M.cropped_extra_info(full, starts, sizes)
        return d


class PosNegCropD(Rnd):
    def __init__(self, keys, label_key, size, n):
        super().__init__()
        self.keys, self.label_key, self.size, self.n = list(keys), label_key, list(size), n

    def __call__(self, d):
        lab = d[self.label_key]
        centers, size = M.pos_neg_crop_centers(self.R, self.size, self.n, 0.5, tuple(lab.shape[1:]),
                                               d["mask_fg_indices"], d["mask_bg_indices"])
        out = []
        for c in centers:
            r = dict(d)
            for k in self.keys:
                r[k] = M.crop_from_center(d[k], c, size)[0]
            out.append(r)
        return out


class CenterCropD:
    def __init__(self, keys, roi):
        self.keys, self.roi = list(keys), list(roi)

    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            d[k] = M.center_spatial_crop(d[k], self.roi)
        return d


class ScaleD(Rnd):
    """RandScaleIntensityd †."""

    def __init__(self, keys, factors):
        super().__init__()
        self.keys, self.f, self.R2 = list(keys), factors, np.random.RandomState()

    def seed(self, s):
        super().seed(s)
        self.R2 = np.random.RandomState(s)
        return self

    def __call__(self, d):
        d = dict(d)
        self.R.rand()            # dict-level gate, prob 1.0
        self.R2.rand()           # array-level gate, prob 1.0
        f = self.R2.uniform(low=-self.f, high=self.f)
        for k in self.keys:
            d[k] = (d[k] * (1 + f)).to(torch.float32)
        return d


class ShiftD(ScaleD):
    """RandShiftIntensityd †."""

    def __call__(self, d):
        d = dict(d)
        self.R.rand()
        self.R2.rand()
        o = self.R2.uniform(low=-self.f, high=self.f)
        for k in self.keys:
            d[k] = (d[k] + o).to(torch.float32)
        return d


class ContrastD(ScaleD):
    """RandAdjustContrastd †: dict-level gate, then the array transform's gate (prob 1.0) and
    ``gamma ~ U(lo, hi)`` on the identically seeded second stream; one gamma for every key."""

    def __init__(self, keys, prob, gamma):
        super().__init__(keys, None)
        self.prob, self.gamma = prob, gamma

    def __call__(self, d):
        d = dict(d)
        if not self.R.rand() < self.prob:
            return d
        self.R2.rand()
        g = self.R2.uniform(low=self.gamma[0], high=self.gamma[1])
        for k in self.keys:
            d[k] = M.adjust_contrast(d[k], g)
        return d


class RicianD(ScaleD):
    """RandRicianNoised †: dict-level gate, then per key the array transform (prob 1.0) drawing its own
    sigma and two normal volumes from the identically seeded second stream."""

    def __init__(self, keys, prob, std):
        super().__init__(keys, std)
        self.prob = prob

    def __call__(self, d):
        d = dict(d)
        if not self.R.rand() < self.prob:
            return d
        for k in self.keys:
            d[k], _, _ = M.rand_rician_noise(d[k], self.R2, self.f)
        return d


class StdShiftD(ScaleD):
    """RandStdShiftIntensityd †."""

    def __init__(self, keys, prob, factors):
        super().__init__(keys, factors)
        self.prob = prob

    def __call__(self, d):
        d = dict(d)
        if not self.R.rand() < self.prob:
            return d
        self.R2.rand()
        f = self.R2.uniform(low=-self.f, high=self.f)
        for k in self.keys:
            d[k] = M.std_shift_intensity(d[k], f)
        return d


class ProbShiftD(ScaleD):
    """RandShiftIntensityd † with a gate probability (the --augment intensity member)."""

    def __init__(self, keys, prob, offsets):
        super().__init__(keys, offsets)
        self.prob = prob

    def __call__(self, d):
        d = dict(d)
        if not self.R.rand() < self.prob:
            return d
        self.R2.rand()
        o = self.R2.uniform(low=-self.f, high=self.f)
        for k in self.keys:
            d[k] = (d[k] + o).to(torch.float32)
        return d


class NoiseD(ScaleD):
    """RandGaussianNoised †: one float64 normal draw of the first key's shape, cast to fp32."""

    def __call__(self, d):
        d = dict(d)
        self.R.rand()
        shape = tuple(d[self.keys[0]].shape)
        _, noise = M.rand_gaussian_noise_draw(self.R2, 1.0, self.f, shape)
        self.last = noise
        for k in self.keys:
            d[k] = d[k].to(torch.float32) + noise
        return d


def _member(keys, name, mult=0.5):
    """modules/augmentations.py:103-186 with max_mult applied."""
    ax = {"x": 0, "y": 1, "z": 2}
    if name == "gaussian_noise":
        return NoiseD(keys, 1 * mult)
    if name == "shift_intensity":
        return ShiftD(keys, 0.5 * mult)
    if name == "scale_intensity":
        return ScaleD(keys, 0.5 * mult)
    if name == "contrast":        # AUG_PARAM_CORRECTION: gamma = 3 * mult + 0.51, a scalar: U(0.5, gamma)
        return ContrastD(keys, 1.0, (0.5, 3 * mult + 0.51))
    if name == "rician_noise":
        return RicianD(keys, 1.0, 0.2 * mult)
    kind, c = name.rsplit("_", 1)
    rng = [0, 0, 0]
    if kind == "rotate":
        a = (np.pi / 6 if c != "z" else np.pi / 16) * mult
        rng[ax[c]] = (-a, a)
        return AffineD(keys, 1.0, ["bilinear"] * len(keys), "zeros", rotate_range=tuple(rng))
    if kind == "translate":
        t = (30 if c != "z" else 5) * mult
        rng[ax[c]] = (-t, t)
        return AffineD(keys, 1.0, ["bilinear"] * len(keys), "zeros", translate_range=tuple(rng))
    if kind == "shear":
        x = 0.5 * mult
        rng[ax[c]] = (1 - x, 1 + x)
        return AffineD(keys, 1.0, ["bilinear"] * len(keys), "zeros", shear_range=tuple(rng))
    if kind == "scale":
        x = 0.3 * mult
        rng[ax[c]] = (1 - x, 1 + x)
        return AffineD(keys, 1.0, ["bilinear"] * len(keys), "zeros", scale_range=tuple(rng))
    raise KeyError(name)


class Workhorse(Rnd):
    """AugmentationWorkhorsed: global ``np.random.choice`` then the members in order."""

    def __init__(self, names, keys, N):
        super().__init__()
        self.names, self.N = list(names), N
        self.members = {n: _member(keys, n) for n in self.names}

    def seed(self, s):
        super().seed(s)
        for n in self.names:
            self.members[n].seed(int(self.R.randint(MAX_SEED, dtype="uint32")))
        return self

    def __call__(self, d):
        for t in np.random.choice(self.names, self.N, replace=False):
            d = self.members[str(t)](d)
        return d


class ConcatD:
    def __init__(self, keys, name):
        self.keys, self.name = list(keys), name

    def __call__(self, d):
        d = dict(d)
        d[self.name] = torch.cat([d[k].to(torch.float32) for k in self.keys], 0)
        return d


# --------------------------------------------------------------------------- builders
def unet(augment, all_keys, image_keys, random_crop_size=None, has_label=True, n_crops=1, flip_axis=(0, 1)):
    modes = ["bilinear" if k in image_keys else "nearest" for k in all_keys]
    prob = 1.0 if "trivial" in augment else 0.2
    aug = [Same()] if "trivial" in augment else []
    if "intensity" in augment:
        aug += [ContrastD(image_keys, prob, (0.5, 1.5)), StdShiftD(image_keys, prob, 0.1)]
    if "affine" in augment:
        aug.append(AffineD(all_keys, prob, modes, rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16]))
    if "shear" in augment:
        aug.append(AffineD(all_keys, prob, modes, shear_range=((0.9, 1.1),) * 3))
    flips = [FlipD(all_keys, 0.25, [a]) for a in flip_axis] if "flip" in augment else []
    chain = Chain([PickOne(aug), *flips]) if "trivial" in augment else Chain([*aug, *flips])
    if random_crop_size is not None:
        pre = [int(i * 1.10) for i in random_crop_size]
        first = PosNegCropD([*image_keys, "mask"], "mask", pre, n_crops) if has_label else RandCropD(image_keys, pre)
        chain = Chain([first, chain, CenterCropD([*image_keys, "mask"] if has_label else image_keys, random_crop_size)])
    return chain


def classification(augment, image_keys, mask_key, flip_axis=(0, 1), prob=0.1, n_transforms_trivial=1):
    keys = list(image_keys) + ([mask_key] if mask_key is not None else [])
    modes = ["bilinear" if k != mask_key else "nearest" for k in keys]
    aug = []
    if "trivial" in augment:
        aug.append(Same())
        prob = 1.0
    if "intensity" in augment:
        aug += [ContrastD(image_keys, prob, (0.5, 1.5)), StdShiftD(image_keys, prob, 0.1), ProbShiftD(image_keys, prob, 0.1)]
    if "flip" in augment:
        combos = [c for i in range(len(flip_axis)) for c in itertools.combinations(flip_axis, i + 1)]
        aug.append(PickOne([FlipD(keys, prob, c) for c in combos]))
    if "affine" in augment:
        aug.append(AffineD(keys, prob, modes, "zeros", translate_range=[4, 4, 1], rotate_range=[np.pi / 16],
                           scale_range=[0.1, 0.1, 0.05]))
    if "shear" in augment:
        aug.append(AffineD(keys, prob, modes, "zeros", shear_range=((0.9, 1.1),) * 3))
    return PickSome(aug, n_transforms_trivial) if "trivial" in augment else Chain(aug)


def ssl(all_keys, copied_keys, roi_size, vicregl, different_crop, names, n_transforms=3):
    roi = tuple(int(x) for x in roi_size)
    if vicregl:
        names = [n for n in names if n.split("_")[0] not in ("rotate", "translate", "shear") and n not in ("scale_x", "scale_y", "scale_z")]
    ts = []
    if vicregl or different_crop:
        ts += [RandCropD(all_keys, roi), RandCropD(copied_keys, roi)]
    else:
        ts.append(RandCropD(list(all_keys) + list(copied_keys), roi))
    ts.append(Workhorse(names, all_keys, n_transforms))
    if copied_keys:
        ts.append(Workhorse(names, copied_keys, n_transforms))
    return ts
