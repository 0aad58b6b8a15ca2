"""CPU ORACLE — test infrastructure only, never the product path.

A restatement, on torch-CPU + numpy, of the arithmetic that adell-mri's
``transform_factory`` chain executes per sample.  The reference only *wires*
MONAI dictionary transforms together (``/root/reference/adell_mri/transform_factory/
augmentations.py:98-176,255-301,427-515`` and ``transforms.py:143-204,430-499,772-820``);
the maths lives in third-party **monai 1.6.0** (``/root/reference/uv.lock:1670-1671``)
which calls ``torch.nn.functional.grid_sample``.  MONAI is not importable in this
image (no wheel, no network), so every function below restates MONAI's published
behaviour († = restated from MONAI 1.3-1.6 upstream, not readable under
/root/reference) and routes the resampling itself through the *same ATen CPU
kernel* MONAI would call (``F.grid_sample``).

PARITY STATUS: **parity unpinned** for resampling / flip / random-crop / noise
numerics — the reference's own tests hold no golden vectors for them
(SURVEY.md §4, §8c).  What *is* pinned: the crop/pad/min-max invariants of
``/root/reference/testing/test_crop_from_maskd.py:166-374`` and the crop counts of
``testing/test_get_all_crops.py:17-33`` (see tests/test_oracle_reference_invariants.py),
and every resampling closed form here is cross-checked against ``F.grid_sample``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""

from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

MAX_SEED = np.iinfo(np.uint32).max + 1  # monai.utils.MAX_SEED †


# --------------------------------------------------------------------------- #
# grid + affine (monai.transforms.utils.create_grid/create_rotate/... †)
# --------------------------------------------------------------------------- #
def create_grid(spatial_size: Sequence[int]) -> torch.Tensor:
    """monai.transforms.utils.create_grid(backend=torch, homogeneous=True) †.

    ``[4, H, W, D]`` fp32: per axis ``linspace(-(d-1)/2, (d-1)/2, d)``, ``ij``
    meshgrid, stacked with a row of ones.
    """
    ranges = [
        torch.linspace(-(d - 1.0) / 2.0, (d - 1.0) / 2.0, int(d), dtype=torch.float32)
        for d in spatial_size
    ]
    coords = torch.meshgrid(*ranges, indexing="ij")
    return torch.stack([*coords, torch.ones_like(coords[0])])


def _sin(th):
    return torch.sin(torch.as_tensor(th, dtype=torch.float32))


def _cos(th):
    return torch.cos(torch.as_tensor(th, dtype=torch.float32))


def create_rotate(radians: Sequence[float]) -> torch.Tensor:
    """3-D ``Rx @ Ry @ Rz`` (leading rotations only if <3 angles) †."""
    radians = list(radians)
    affine = None
    if len(radians) >= 1:
        s, c = _sin(radians[0]), _cos(radians[0])
        affine = torch.eye(4)
        affine[1, 1], affine[1, 2] = c, -s
        affine[2, 1], affine[2, 2] = s, c
    if len(radians) >= 2:
        s, c = _sin(radians[1]), _cos(radians[1])
        m = torch.eye(4)
        m[0, 0], m[0, 2] = c, s
        m[2, 0], m[2, 2] = -s, c
        affine = affine @ m
    if len(radians) >= 3:
        s, c = _sin(radians[2]), _cos(radians[2])
        m = torch.eye(4)
        m[0, 0], m[0, 1] = c, -s
        m[1, 0], m[1, 1] = s, c
        affine = affine @ m
    if affine is None:
        raise ValueError("radians must be non empty.")
    return affine


def create_shear(coefs: Sequence[float]) -> torch.Tensor:
    """3-D shear, coefficients (a01,a02,a10,a12,a20,a21) zero-padded to 6 †."""
    coefs = list(coefs)[:6] + [0.0] * max(0, 6 - len(coefs))
    out = torch.eye(4)
    out[0, 1], out[0, 2] = coefs[0], coefs[1]
    out[1, 0], out[1, 2] = coefs[2], coefs[3]
    out[2, 0], out[2, 1] = coefs[4], coefs[5]
    return out


def create_translate(shift: Sequence[float]) -> torch.Tensor:
    out = torch.eye(4)
    for i, a in enumerate(list(shift)[:3]):
        out[i, 3] = a
    return out


def create_scale(factors: Sequence[float]) -> torch.Tensor:
    f = list(factors)[:3] + [1.0] * max(0, 3 - len(factors))
    return torch.diag(torch.as_tensor(f + [1.0], dtype=torch.float32))


def compose_affine(rotate=None, shear=None, translate=None, scale=None) -> torch.Tensor:
    """monai AffineGrid matrix: ``eye @ rotate @ shear @ translate @ scale`` (fp32) †."""
    affine = torch.eye(4)
    if rotate:
        affine = affine @ create_rotate(rotate)
    if shear:
        affine = affine @ create_shear(shear)
    if translate:
        affine = affine @ create_translate(translate)
    if scale:
        affine = affine @ create_scale(scale)
    return affine


def affine_grid(affine: torch.Tensor, spatial_size: Sequence[int]) -> torch.Tensor:
    """``(affine @ grid.view(4, -1)).view(4, H, W, D)`` — the literal MONAI product †.

    On this image torch routes the [4,4]@[4,N] product to MKL sgemm, whose result
    equals an fp32 FMA chain in k order (checked in tests/test_oracle_c_restatement.py).
    """
    grid = create_grid(spatial_size)
    affine = affine.to(torch.float32)
    return (affine @ grid.view(4, -1)).view(4, *grid.shape[1:])


def resample(img: torch.Tensor, grid: torch.Tensor, mode: str, padding_mode: str) -> torch.Tensor:
    """monai.transforms.Resample.__call__ (torch backend, norm_coords, align_corners=False) †.

    ``img`` is ``[C, H, W, D]``; ``grid`` is ``[>=3, h, w, d]`` in centred voxel units.
    """
    img_t = img.to(torch.float32)
    sr = 3
    grid_t = torch.movedim(grid[list(range(sr - 1, -1, -1))], 0, -1).unsqueeze(0)
    grid_t = grid_t.clone(memory_format=torch.contiguous_format)
    for i, dim in enumerate(img_t.shape[sr:0:-1]):
        grid_t[0, ..., i] *= 2.0 / max(2, dim)
    out = F.grid_sample(
        img_t.unsqueeze(0), grid_t, mode=mode, padding_mode=padding_mode, align_corners=False
    )[0]
    return out


def affine_resample(
    img: torch.Tensor,
    affine: torch.Tensor,
    mode: str = "bilinear",
    padding_mode: str = "reflection",
    spatial_size: Sequence[int] | None = None,
) -> torch.Tensor:
    """One RandAffine(d) firing for one key: grid from ``spatial_size`` (default: image size)."""
    sp = tuple(img.shape[1:]) if spatial_size is None else tuple(spatial_size)
    return resample(img, affine_grid(affine, sp), mode, padding_mode)


# --------------------------------------------------------------------------- #
# random-parameter draws (RandomState order) †
# --------------------------------------------------------------------------- #
def _get_rand_param(R: np.random.RandomState, param_range, add_scalar: float = 0.0):
    out = []
    for f in param_range or []:
        if isinstance(f, (list, tuple, np.ndarray)):
            if len(f) != 2:
                raise ValueError("If giving range as [min,max], should only have two elements per dim.")
            out.append(R.uniform(f[0], f[1]) + add_scalar)
        elif f is not None:
            out.append(R.uniform(-f, f) + add_scalar)
    return out


def rand_affine_grid_params(R, rotate_range=None, shear_range=None, translate_range=None, scale_range=None):
    """RandAffineGrid.randomize: rotate, shear, translate, scale(+1.0) in that order †."""
    return dict(
        rotate=_get_rand_param(R, rotate_range),
        shear=_get_rand_param(R, shear_range),
        translate=_get_rand_param(R, translate_range),
        scale=_get_rand_param(R, scale_range, 1.0),
    )


class RandAffinedDraws:
    """Stream-faithful (best effort, ††) replica of the draws in RandAffined.__call__.

    Three RandomStates seeded identically (``set_random_state`` propagates the seed
    to RandAffined, its RandAffine and its RandAffineGrid).  Per call:
      outer.rand() < prob; inner.rand() (prob 1.0); grid params (discarded);
      if fired: grid params again (USED — RandAffineGrid.__call__ re-randomises);
      then per key, ONLY when resampling (`if do_resampling:` in RandAffined.__call__, MONAI >= 1.0; with
      spatial_size=None that is `fired`): inner.rand() + grid params (discarded).
    ``idle_per_key_draws=True`` restores the round-1 reading (per-key draws on idle calls as well).
    """

    def __init__(self, prob, rotate_range=None, shear_range=None, translate_range=None, scale_range=None, n_keys=1,
                 idle_per_key_draws=False):
        self.idle_per_key_draws = idle_per_key_draws
        self.prob = prob
        self.ranges = dict(
            rotate_range=rotate_range, shear_range=shear_range, translate_range=translate_range, scale_range=scale_range
        )
        self.n_keys = n_keys
        self.set_random_state(None)

    def set_random_state(self, seed):
        self.R = np.random.RandomState(seed)
        self.R_inner = np.random.RandomState(seed)
        self.R_grid = np.random.RandomState(seed)
        return self

    def draw(self):
        do = self.R.rand() < self.prob
        self.R_inner.rand()
        rand_affine_grid_params(self.R_grid, **self.ranges)
        used = None
        if do:
            used = rand_affine_grid_params(self.R_grid, **self.ranges)
        for _ in range(self.n_keys if (do or self.idle_per_key_draws) else 0):
            self.R_inner.rand()
            rand_affine_grid_params(self.R_grid, **self.ranges)
        return do, used


def compose_set_random_state(seed: int, n_randomizable: int) -> list[int]:
    """Compose.set_random_state(seed): child seeds ``R.randint(MAX_SEED, dtype='uint32')`` †."""
    R = np.random.RandomState(seed)
    return [int(R.randint(MAX_SEED, dtype="uint32")) for _ in range(n_randomizable)]


# --------------------------------------------------------------------------- #
# flip / pad / crop (integer, exact) †
# --------------------------------------------------------------------------- #
def flip(img: torch.Tensor, spatial_axis) -> torch.Tensor:
    """monai Flip: ``torch.flip(img, dims=[a+1 ...])``."""
    axes = [spatial_axis] if isinstance(spatial_axis, int) else list(spatial_axis)
    return torch.flip(img, [a + 1 for a in axes])


def spatial_pad(img: torch.Tensor, spatial_size: Sequence[int]) -> torch.Tensor:
    """SpatialPad(method='symmetric', mode='constant', value 0) †: before=⌊w/2⌋, after=w-⌊w/2⌋."""
    pads = []
    for d, s in zip(img.shape[1:], spatial_size):
        w = max(int(s) - int(d), 0) if s is not None and s > 0 else 0
        pads.append((w // 2, w - w // 2))
    flat = []
    for lo, hi in reversed(pads):
        flat += [lo, hi]
    return F.pad(img, flat, mode="constant", value=0.0)


def center_crop_slices(shape: Sequence[int], roi: Sequence[int]):
    """CenterSpatialCrop †: centre=⌊size/2⌋, start=max(centre-⌊roi/2⌋,0), end=start+roi (clipped)."""
    out = []
    for d, r in zip(shape, roi):
        r = d if (r is None or r <= 0) else min(int(r), d)
        start = max(d // 2 - r // 2, 0)
        end = min(start + r, d)
        out.append(slice(start, end))
    return out


def center_spatial_crop(img: torch.Tensor, roi: Sequence[int]) -> torch.Tensor:
    sl = center_crop_slices(img.shape[1:], roi)
    return img[(slice(None), *sl)]


def rand_spatial_crop_draw(R, shape, roi_size, random_size=False, max_roi_size=None):
    """RandSpatialCrop.randomize † — returns (starts, sizes)."""
    size = [d if (r is None or r <= 0) else min(int(r), d) for d, r in zip(shape, roi_size)]
    if random_size:
        mx = list(shape) if max_roi_size is None else [min(int(m), d) for m, d in zip(max_roi_size, shape)]
        size = [int(R.randint(low=size[i], high=mx[i] + 1)) for i in range(len(shape))]
    valid = [d - s + 1 for d, s in zip(shape, size)]
    starts = [int(R.randint(v)) for v in valid]
    return starts, size


def crop(img: torch.Tensor, starts, sizes) -> torch.Tensor:
    sl = [slice(s, s + z) for s, z in zip(starts, sizes)]
    return img[(slice(None), *sl)]


def cropped_extra_info(shape, starts, sizes):
    """``extra_info['cropped'] = [start0, size0-end0, start1, ...]`` († consumed at
    /root/reference/adell_mri/transform_factory/augmentations.py:402-406,460-473)."""
    out = []
    for d, s, z in zip(shape, starts, sizes):
        out += [int(s), int(d - (s + z))]
    return out


def flatten_box(box, roi_size):
    """/root/reference/adell_mri/transform_factory/augmentations.py:402-406."""
    box1 = np.array(box[::2])
    box2 = np.array(roi_size) - np.array(box[1::2])
    return np.concatenate([box1, box2]).astype(np.float32)


def pos_neg_crop_centers(R, spatial_size, num_samples, pos_ratio, label_shape, fg_indices, bg_indices):
    """monai generate_pos_neg_label_crop_centers + correct_crop_centers(allow_smaller=True) †."""
    spatial_size = [min(int(s), int(d)) for s, d in zip(spatial_size, label_shape)]
    centers = []
    fg_indices = np.asarray(fg_indices)
    bg_indices = np.asarray(bg_indices)
    if len(fg_indices) == 0 or len(bg_indices) == 0:
        pos_ratio = 0 if len(fg_indices) == 0 else 1
    for _ in range(num_samples):
        indices_to_use = fg_indices if R.rand() < pos_ratio else bg_indices
        random_int = R.randint(len(indices_to_use))
        idx = indices_to_use[random_int]
        center = np.unravel_index(idx, label_shape)
        # correct_crop_centers
        valid_start = np.floor_divide(spatial_size, 2)
        valid_end = np.subtract(np.array(label_shape) + 1, np.array(spatial_size) / 2).astype(np.uint16)
        for i, vs in enumerate(valid_start):
            if vs == valid_end[i]:
                valid_end[i] += 1
        c = []
        for ci, vs, ve in zip(center, valid_start, valid_end):
            c.append(int(min(max(ci, vs), ve - 1)))
        centers.append(c)
    return centers, spatial_size


def crop_from_center(img: torch.Tensor, center, roi):
    """SpatialCrop(roi_center, roi_size) †: start=max(c-⌊r/2⌋,0), end=max(start+r, start)."""
    starts = [max(int(c) - int(r) // 2, 0) for c, r in zip(center, roi)]
    return crop(img, starts, [int(r) for r in roi]), starts


# --------------------------------------------------------------------------- #
# intensity †
# --------------------------------------------------------------------------- #
def scale_intensity(img: torch.Tensor, minv=0.0, maxv=1.0, factor=None) -> torch.Tensor:
    """monai ScaleIntensity (channel_wise=False) → rescale_array †.

    minv/maxv form: ``(x-min)/(max-min)*(maxv-minv)+minv``; constant image ⇒ ``x*minv``.
    factor form: ``x*(1+factor)``.
    """
    img = img.to(torch.float32)
    if minv is not None or maxv is not None:
        mina, maxa = img.min(), img.max()
        if mina == maxa:
            return img * minv if minv is not None else img
        norm = (img - mina) / (maxa - mina)
        if minv is None or maxv is None:
            return norm
        return (norm * (maxv - minv)) + minv
    if factor is not None:
        return (img * (1 + factor)).to(torch.float32)
    raise ValueError("Incompatible values: minv=None or maxv=None and factor=None.")


def conditional_rescaling(img: torch.Tensor, max_value: float, scale: float) -> torch.Tensor:
    """/root/reference/adell_mri/utils/monai_transforms/image_intensity_ops.py:71-74."""
    if img.max() > max_value:
        img = img * scale
    return img


def offset(img: torch.Tensor, offset=None) -> torch.Tensor:
    """/root/reference/adell_mri/utils/monai_transforms/image_intensity_ops.py:119-121."""
    o = img.min() if offset is None else offset
    return img - o


def normalize_intensity(img: torch.Tensor, nonzero=False, channel_wise=False) -> torch.Tensor:
    """monai NormalizeIntensity †: (x-mean)/std, population std, std==0 ⇒ 1."""

    def _norm(x):
        x = x.to(torch.float32)
        sl = (x != 0) if nonzero else torch.ones_like(x, dtype=torch.bool)
        if not sl.any():
            return x
        _sub = x[sl].mean()
        _div = x[sl].std(unbiased=False)
        if _div == 0.0:
            _div = torch.tensor(1.0)
        out = x.clone()
        out[sl] = (x[sl] - _sub) / _div
        return out

    if channel_wise:
        return torch.stack([_norm(c) for c in img])
    return _norm(img)


def percentile(x: torch.Tensor, q: float) -> float:
    """monai.transforms.utils_pytorch_numpy_unification.percentile †:
    >1e6 elements ⇒ numpy 'linear' percentile (float64 lerp, cast back to the
    tensor dtype); else torch.quantile."""
    if x.numel() > 1_000_000:
        # MONAI hands q over as a float64 ndarray -> numpy interpolates in float64
        r = np.percentile(x.numpy(), np.asarray(q, dtype=np.float64))
        return torch.as_tensor(r).to(x.dtype)
    return torch.quantile(x, torch.as_tensor(q / 100.0, dtype=x.dtype))


def scale_intensity_range(img, a_min, a_max, b_min=None, b_max=None, clip=False):
    img = img.to(torch.float32)
    if a_max - a_min == 0.0:
        if b_min is None:
            return img - a_min
        return img - a_min + b_min
    img = (img - a_min) / (a_max - a_min)
    if b_min is not None and b_max is not None:
        img = img * (b_max - b_min) + b_min
    if clip:
        img = torch.clamp(img, b_min, b_max)
    return img


def scale_intensity_range_percentiles(img, lower, upper, b_min, b_max, clip=False):
    """monai ScaleIntensityRangePercentiles (relative=False, channel_wise=False) †."""
    a_min = percentile(img.to(torch.float32), lower)
    a_max = percentile(img.to(torch.float32), upper)
    return scale_intensity_range(img, a_min, a_max, b_min, b_max, clip)


def rand_gaussian_noise_draw(R, prob, std, shape, mean=0.0, sample_std=True):
    """RandGaussianNoise.randomize †: rand()<prob; σ=uniform(0,std); normal(mean,σ,size) f64→f32."""
    do = R.rand() < prob
    if not do:
        return False, None
    s = R.uniform(0, std) if sample_std else std
    noise = R.normal(mean, s, size=shape)
    return True, torch.as_tensor(noise.astype(np.float32))


# --------------------------------------------------------------------------- #
# collation (/root/reference/adell_mri/utils/utils.py:308-377)
# --------------------------------------------------------------------------- #
def safe_collate(X):
    def cat(x):
        try:
            x = [torch.as_tensor(y) for y in x]
        except Exception:
            return x
        try:
            return torch.stack(x)
        except Exception:
            return x

    example = X[0]
    if isinstance(example, list):
        return [cat(e) for e in zip(*X)]
    out = {}
    for k in example.keys():
        out[k] = cat([x[k] if k in x else None for x in X])
    return out


def safe_collate_crops(X):
    flat = []
    for x in X:
        for xx in x:
            flat.append(xx)
    return safe_collate(flat)


# --------------------------------------------------------------------------- #
# canonical item semantics (what ONE fused-gather item must reproduce)
# --------------------------------------------------------------------------- #
def canonical_item(
    parent: torch.Tensor,
    *,
    pre_ops: Sequence[tuple] = (),
    affine: torch.Tensor | None = None,
    mode: str = "bilinear",
    padding_mode: str = "zeros",
    post_ops: Sequence[tuple] = (),
    noise: torch.Tensor | None = None,
    post_scale: float | None = None,
    post_offset: float | None = None,
) -> torch.Tensor:
    """Execute the reference op sequence for one key, one op at a time, each op
    materialising a tensor exactly as MONAI would.  ``pre_ops``/``post_ops`` are
    tuples ``("flip", axes) | ("pad", size) | ("center_crop", roi) | ("crop", starts, sizes)``.
    """

    def run(img, ops):
        for op in ops:
            if op[0] == "flip":
                img = flip(img, op[1])
            elif op[0] == "pad":
                img = spatial_pad(img, op[1])
            elif op[0] == "center_crop":
                img = center_spatial_crop(img, op[1])
            elif op[0] == "crop":
                img = crop(img, op[1], op[2])
            else:
                raise ValueError(op[0])
        return img

    img = run(parent.to(torch.float32), pre_ops)
    if affine is not None:
        img = affine_resample(img, affine, mode, padding_mode)
    img = run(img, post_ops)
    if post_scale is not None:
        img = img * torch.tensor(post_scale, dtype=torch.float32)
    if post_offset is not None:
        img = img + torch.tensor(post_offset, dtype=torch.float32)
    if noise is not None:
        img = img + noise
    return img.contiguous()


def all_flip_combinations(flip_axis):
    """/root/reference/adell_mri/transform_factory/augmentations.py:255-269."""
    out = []
    for i in range(len(flip_axis)):
        out.extend(itertools.combinations(flip_axis, i + 1))
    return out


# --------------------------------------------------------------------------- --augment intensity
def adjust_contrast(img: torch.Tensor, gamma: float) -> torch.Tensor:
    """monai AdjustContrast.__call__ † (invert_image=False, retain_stats=False):
    ``((img - min) / float(range + 1e-7)) ** gamma * range + min`` over the whole array
    (/root/reference/adell_mri/transform_factory/augmentations.py:66-76,219-232)."""
    img = img.to(torch.float32)
    epsilon = 1e-7
    img_min = img.min()
    img_range = img.max() - img_min
    return ((img - img_min) / float(img_range + epsilon)) ** gamma * img_range + img_min


def rand_rician_noise(img: torch.Tensor, R, std: float, mean: float = 0.0, sample_std: bool = True):
    """monai RandRicianNoise.__call__ † (prob 1.0, channel_wise=False, relative=False): its own
    ``R.rand()`` gate, ``sigma = R.uniform(0, std)``, two ``R.normal(mean, sigma, size=img.shape)``
    volumes cast to the image dtype, ``sqrt((img + n1)**2 + n2**2)``
    (/root/reference/adell_mri/modules/augmentations.py:53,86,117).  Returns (out, n1, n2).
    torch's CPU sqrt is MKL VML's (within 1 ulp of the IEEE square root, not always equal to it)."""
    img = img.to(torch.float32)
    R.rand()
    s = R.uniform(0, std) if sample_std else std
    n1 = torch.as_tensor(R.normal(mean, s, size=tuple(img.shape)).astype(np.float32))
    n2 = torch.as_tensor(R.normal(mean, s, size=tuple(img.shape)).astype(np.float32))
    return torch.sqrt((img + n1) ** 2 + n2 ** 2), n1, n2


def resized(img: torch.Tensor, spatial_size: Sequence[int], mode: str = "area") -> torch.Tensor:
    """monai Resize.__call__ † (size_mode="all", anti_aliasing=False): a no-op when the shape already
    matches, else ``F.interpolate(img[None].float(), size, mode)`` — "area" is ATen's
    adaptive_avg_pool3d (/root/reference/adell_mri/transform_factory/augmentations.py:427-444 uses MONAI's
    default mode "area"; transforms.py:157-167 passes "area" / "nearest" per key)."""
    size = tuple(int(x) for x in spatial_size)
    if tuple(img.shape[1:]) == size:
        return img
    return torch.nn.functional.interpolate(img[None].to(torch.float32), size=size, mode=mode)[0]


def std_shift_intensity(img: torch.Tensor, factor: float) -> torch.Tensor:
    """monai StdShiftIntensity._stdshift † (nonzero=False, channel_wise=False): ``img + factor *
    std(img)`` with the population standard deviation of the whole array."""
    img = img.to(torch.float32)
    offset = factor * torch.std(img, unbiased=False)
    return img + offset


def spacing(img: torch.Tensor, spacing_in: Sequence[float], pixdim: Sequence[float], mode: str = "bilinear",
            padding_mode: str = "border", dtype=torch.float64) -> torch.Tensor:
    """monai.transforms.Spacing † on an axis-aligned volume (affine = diag(spacing_in), origin anywhere):
    ``new_affine = zoom_affine(affine, pixdim)`` keeps the origin; ``compute_shape_offset`` sizes the output as
    ``round((S-1) * spacing_in / pixdim + 1)``; ``SpatialResample(align_corners=False)`` maps output index ``o`` to
    source index ``solve(src_affine, dst_affine) @ o = o * pixdim / spacing_in``; ``grid_sample`` evaluates it in
    ``dtype`` (MONAI's default float64) and the result is cast back to float32.
    (/root/reference/adell_mri/transform_factory/transforms.py:133-140,444-454)"""
    S = np.asarray(img.shape[1:], np.float64)
    r = np.asarray(pixdim, np.float64) / np.asarray(spacing_in, np.float64)
    out = np.round((S - 1.0) / r + 1.0).astype(np.int64)
    x = img.to(dtype)[None]
    axes = []
    for a in range(3):
        u = torch.arange(int(out[a]), dtype=dtype) * float(r[a])           # source index of output voxel o
        axes.append((2.0 * u + 1.0) / float(S[a]) - 1.0)                    # grid_sample, align_corners=False
    g0, g1, g2 = torch.meshgrid(*axes, indexing="ij")
    grid = torch.stack([g2, g1, g0], dim=-1)[None]                          # grid_sample wants (x, y, z) = axes (2, 1, 0)
    y = F.grid_sample(x, grid, mode=mode, padding_mode=padding_mode, align_corners=False)
    return y[0].to(torch.float32)
