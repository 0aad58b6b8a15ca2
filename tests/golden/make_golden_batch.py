"""Generates tests/golden/batch_preprocessing.npz by running the REFERENCE ITSELF
(/root/reference/adell_mri/utils/batch_preprocessing.py) in the build container.  The module is
loaded from its file (its package __init__ imports monai, which is not installed); its only
intra-package import, the logger factory, is stubbed with the standard library's.

    python tests/golden/make_golden_batch.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/adell_mri/utils/batch_preprocessing.py"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, shape, label dtype, kwargs, number of consecutive calls on the same object
    ("mixup_all", (6, 2, 5, 4, 3), "float32", dict(mixup_alpha=0.4, seed=42), 2),
    ("partial", (7, 1, 6, 5, 4), "float32", dict(mixup_alpha=0.3, partial_mixup=0.5, seed=7), 3),
    ("smooth_then_partial_int_labels", (5, 3, 4, 4, 2), "int64", dict(label_smoothing=0.1, mixup_alpha=1.0, partial_mixup=0.6, seed=3), 2),
    ("smooth_only", (4, 1, 3, 3, 3), "float32", dict(label_smoothing=0.2), 1),
    ("mixup_int_labels_truncate", (6, 1, 4, 3, 5), "int64", dict(mixup_alpha=0.5, seed=11), 1),
]


def load_reference():
    for name in ("adell_mri", "adell_mri.utils"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    pl = types.ModuleType("adell_mri.utils.python_logging")
    pl.get_logger = lambda name: logging.getLogger(name)
    sys.modules["adell_mri.utils.python_logging"] = pl
    spec = importlib.util.spec_from_file_location("reference_batch_preprocessing", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs(name, shape, ydtype, call):
    R = np.random.RandomState(sum(map(ord, name)) * 31 + call)
    x = R.rand(*shape).astype(np.float32) * 3 - 1
    y = (R.rand(shape[0]) > 0.5).astype(ydtype)
    return x, y


def main():
    ref = load_reference()
    out = {}
    for name, shape, ydtype, kw, calls in CASES:
        bp = ref.BatchPreprocessing(**kw)
        for c in range(calls):
            x, y = inputs(name, shape, ydtype, c)
            X, Y = bp(torch.from_numpy(x.copy()), torch.from_numpy(y.copy()))
            out[f"{name}/{c}/x"] = X.numpy()
            out[f"{name}/{c}/y"] = Y.numpy()
    np.savez_compressed(os.path.join(HERE, "batch_preprocessing.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
