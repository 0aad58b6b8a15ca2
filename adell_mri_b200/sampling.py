"""Host-side random-parameter draws that follow the reference's RandomState streams.

The reference seeds its pipelines with ``Compose.set_random_state(seed)``
(/root/reference/adell_mri/entrypoints/segmentation/train.py:449, classification/train.py:243,
ssl/train_3d.py:196) and MONAI then draws, per sample, in a fixed order from per-transform
``numpy.random.RandomState`` objects.  The classes below reproduce those draw orders († —
restated from MONAI 1.3-1.6; MONAI is not importable here) so that, given the same seed,
the same geometric parameters reach the GPU.  Parity of *voxels* is asserted at the
parameter boundary; stream fidelity is best effort and documented as such in DESIGN.md.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

MAX_SEED = np.iinfo(np.uint32).max + 1


def _issequence(f):
    return isinstance(f, (list, tuple, np.ndarray))


def rand_param(R: np.random.RandomState, param_range, add_scalar: float = 0.0) -> list[float]:
    """RandAffineGrid._get_rand_param †."""
    out = []
    for f in param_range or []:
        if _issequence(f):
            if len(f) != 2:
                raise ValueError("If giving range as [min,max], should only have two elements per dim.")
            out.append(R.uniform(f[0], f[1]) + add_scalar)
        elif f is not None:
            out.append(R.uniform(-f, f) + add_scalar)
    return out


#: † MONAI's ``RandAffined.__call__`` re-randomises its inner ``RandAffine`` once per key through
#: ``self.rand_affine(d[key], None, mode, padding_mode, True, grid)`` — but only inside ``if do_resampling:``
#: (MONAI >= 1.0; with ``spatial_size=None``, as the reference always uses, ``do_resampling == fired``).  A call
#: that does not fire therefore consumes one outer gate draw, one inner gate draw and ONE grid randomize, nothing
#: per key.  Round 1 consumed the per-key draws on idle calls too; that reading is kept behind this switch so a
#: maintainer with MONAI installed can compare both against the real stream.
PER_KEY_DRAWS_WHEN_IDLE = False


class RandAffineSampler:
    """Draw order of ``monai.transforms.RandAffined.__call__`` (three identically seeded
    streams: the dict transform, its RandAffine and its RandAffineGrid) †.

    ``per_key_draws_when_idle``: whether a call whose gate did NOT fire still consumes the per-key
    ``RandAffine.randomize()`` draws (``None`` = module default :data:`PER_KEY_DRAWS_WHEN_IDLE`, i.e. the
    MONAI >= 1.0 reading: it does not)."""

    def __init__(self, prob=0.1, rotate_range=None, shear_range=None, translate_range=None, scale_range=None,
                 per_key_draws_when_idle: bool | None = None):
        self.prob = prob
        self.per_key_draws_when_idle = PER_KEY_DRAWS_WHEN_IDLE if per_key_draws_when_idle is None else bool(per_key_draws_when_idle)
        self.rotate_range, self.shear_range = rotate_range, shear_range
        self.translate_range, self.scale_range = translate_range, scale_range
        self.set_random_state()

    def set_random_state(self, seed=None, state=None):
        if state is not None:
            self.R = self.R_inner = self.R_grid = state  # a shared state object, as MONAI does
        else:
            self.R = np.random.RandomState(seed)
            self.R_inner = np.random.RandomState(seed)
            self.R_grid = np.random.RandomState(seed)
        return self

    def _grid_params(self):
        R = self.R_grid
        return dict(
            rotate=rand_param(R, self.rotate_range),
            shear=rand_param(R, self.shear_range),
            translate=rand_param(R, self.translate_range),
            scale=rand_param(R, self.scale_range, 1.0),
        )

    def draw(self, n_keys: int = 1):
        """One ``__call__``: returns ``(fired, params-or-None)``."""
        fired = self.R.rand() < self.prob          # RandAffined.randomize
        self.R_inner.rand()                        # RandAffine.randomize (prob=1.0) ...
        self._grid_params()                        # ... -> RandAffineGrid.randomize (discarded)
        used = self._grid_params() if fired else None  # RandAffineGrid.__call__ re-randomises: USED
        if fired or self.per_key_draws_when_idle:
            for _ in range(n_keys):                # per key: RandAffine.__call__(randomize=True), under `if do_resampling`
                self.R_inner.rand()
                self._grid_params()
        return fired, used


    # ---- vectorised over a batch of consecutive calls (bit-identical to calling draw() B times) ----
    def _n_per_randomize(self):
        return sum(len(r) if r is not None else 0 for r in (self.rotate_range, self.shear_range,
                                                            self.translate_range, self.scale_range))

    def _params_from_uniforms(self, u: np.ndarray):
        """``u``: [m, K] raw ``random_sample`` values of m RandAffineGrid.randomize calls -> dict of arrays."""
        out, col = {}, 0
        for name, rng, add in (("rotate", self.rotate_range, 0.0), ("shear", self.shear_range, 0.0),
                               ("translate", self.translate_range, 0.0), ("scale", self.scale_range, 1.0)):
            cols = []
            for f in rng or []:
                lo, hi = (f[0], f[1]) if _issequence(f) else (-f, f)
                cols.append(lo + (hi - lo) * u[:, col] + add)  # numpy's uniform(): low + (high-low)*random_sample()
                col += 1
            out[name] = np.stack(cols, axis=1) if cols else np.zeros((u.shape[0], 0))
        return out

    def draw_batch(self, batch: int, n_keys: int = 1):
        """``batch`` consecutive ``__call__``s at once: returns ``(fired[batch], params)`` with
        ``params[name]`` of shape ``[n_fired, k]`` — the same values, consumed from the same three
        streams in the same order, as ``batch`` calls of :meth:`draw`."""
        K = self._n_per_randomize()
        if self.R is self.R_inner:  # shared state object: the streams interleave, no batching possible
            res = [self.draw(n_keys) for _ in range(batch)]
            fired = np.array([r[0] for r in res], bool)
            plist = [r[1] for r in res if r[0]]
            names = ("rotate", "shear", "translate", "scale")
            return fired, {k: np.asarray([p[k] for p in plist], np.float64).reshape(len(plist), -1) for k in names}
        fired = self.R.random_sample(batch) < self.prob
        if fired.all():   # (every call fires, e.g. the workhorse members with prob 1: a fixed stride through the streams)
            calls = 2 + n_keys
            self.R_inner.random_sample(batch * (1 + n_keys))
            u = self.R_grid.random_sample(batch * calls * K).reshape(batch, calls, K) if K else np.zeros((batch, calls, 0))
            return fired, self._params_from_uniforms(u[:, 1])
        per_key = np.full(batch, n_keys, np.int64) if self.per_key_draws_when_idle else n_keys * fired.astype(np.int64)
        self.R_inner.random_sample(batch + int(per_key.sum()))
        calls = 1 + fired.astype(np.int64) + per_key         # randomize() calls on the grid stream per sample
        start = np.concatenate([[0], np.cumsum(calls)[:-1]])
        u = self.R_grid.random_sample(int(calls.sum()) * K).reshape(-1, K) if K else np.zeros((int(calls.sum()), 0))
        used = u[start[fired] + 1]                             # the second randomize of a firing call is the used one
        return fired, self._params_from_uniforms(used)


def draw_used_uniforms(smp: RandAffineSampler, m: int, n_keys: int) -> np.ndarray:
    """``m`` consecutive calls of a sampler that ALWAYS fires (prob >= 1, e.g. the workhorse members): the raw
    uniforms ``[m, K]`` of the USED randomize of each call, the three streams advanced exactly as ``m`` calls of
    :meth:`RandAffineSampler.draw` would (one numpy call per stream)."""
    K = smp._n_per_randomize()
    calls = 2 + n_keys
    smp.R.random_sample(m)
    smp.R_inner.random_sample(m * (1 + n_keys))
    return smp.R_grid.random_sample(m * calls * K).reshape(m, calls, K)[:, 1] if K else np.zeros((m, 0))


def child_seeds(seed: int, n: int) -> list[int]:
    """``Compose.set_random_state(seed)``: one ``R.randint(MAX_SEED, dtype=uint32)`` per Randomizable child †."""
    R = np.random.RandomState(seed)
    return [int(R.randint(MAX_SEED, dtype="uint32")) for _ in range(n)]
