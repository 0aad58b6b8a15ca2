"""`safe_collate` / `safe_collate_crops`: golden outputs produced by executing the reference's own
function definitions (tests/golden/make_golden_collate.py); the product's collate functions must
return the same structure and values on plain tensors (stack when shapes agree, the list when they
do not or a key is missing, crops flattened first).  Pending entries are covered by
tests/test_lazy_pipelines.py."""

import importlib.util
import os

import numpy as np

from adell_mri_b200 import collate

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_collate", os.path.join(HERE, "golden", "make_golden_collate.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "collate.npz"))


def test_collate_matches_reference_functions():
    fns = {"safe_collate": collate.safe_collate, "safe_collate_crops": collate.safe_collate_crops}
    store = {}
    for name, (fn, x) in G.cases().items():
        G.flatten(name, fns[fn](x), store)
    assert sorted(store) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(store[k], GOLD[k]), k
