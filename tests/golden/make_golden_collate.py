"""Generates tests/golden/collate.npz by executing the REFERENCE's own `safe_collate`,
`safe_collate_crops` and `unpack_crops` (/root/reference/adell_mri/utils/utils.py:230-244,308-377).
The module itself imports half of the package (losses, MONAI, SimpleITK), so the three function
definitions are taken from its source with `ast` and executed unchanged in a namespace that only
provides `torch` and the type aliases of their annotations.

    python tests/golden/make_golden_collate.py        # needs /root/reference; not run on the GPU box
"""
import ast
import os

import numpy as np
import torch

SRC = "/root/reference/adell_mri/utils/utils.py"
HERE = os.path.dirname(os.path.abspath(__file__))
WANTED = ("unpack_crops", "safe_collate", "safe_collate_crops")


def load_reference_functions():
    tree = ast.parse(open(SRC).read(), SRC)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert sorted(n.name for n in body) == sorted(WANTED)
    ns = {"torch": torch, "np": np, "TensorIterable": object, "TensorList": object}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return {k: ns[k] for k in WANTED}


def cases():
    """name -> (function name, input).  Plain CPU tensors / scalars, as DataLoader workers hand them over."""
    R = np.random.RandomState(0)
    t = lambda *s: torch.from_numpy(R.rand(*s).astype(np.float32))
    same = [{"image": t(3, 6, 5, 4), "mask": t(1, 6, 5, 4), "label": 1} for _ in range(4)]
    ragged = [{"image": t(1, 6, 5, 4), "box": t(2 + i, 6)} for i in range(3)]
    missing = [{"image": t(1, 4, 4, 2), "extra": t(2)}, {"image": t(1, 4, 4, 2)}]
    lists = [[t(2, 3), t(4)], [t(2, 3), t(4)], [t(2, 3), t(4)]]
    crops = [[{"image": t(1, 4, 4, 2), "mask": t(1, 4, 4, 2)} for _ in range(2)] for _ in range(3)]
    return {"same_shapes": ("safe_collate", same), "ragged_key": ("safe_collate", ragged), "missing_key": ("safe_collate", missing),
            "list_samples": ("safe_collate", lists), "crops": ("safe_collate_crops", crops)}


def flatten(prefix, out, store):
    """Record structure + content: tensors as arrays, lists as `<name>/list/<i>`, None as a marker."""
    if isinstance(out, dict):
        store[prefix + "/keys"] = np.array(list(out.keys()))
        for k, v in out.items():
            flatten(f"{prefix}/{k}", v, store)
    elif isinstance(out, (list, tuple)):
        store[prefix + "/len"] = np.array(len(out))
        for i, v in enumerate(out):
            flatten(f"{prefix}/{i}", v, store)
    elif out is None:
        store[prefix + "/none"] = np.array(1)
    else:
        store[prefix + "/tensor"] = np.asarray(out)


def main():
    fns = load_reference_functions()
    store = {}
    for name, (fn, x) in cases().items():
        flatten(name, fns[fn](x), store)
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **store)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()
