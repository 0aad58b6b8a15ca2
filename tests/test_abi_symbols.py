"""CPU: the C-ABI library loads and exports every symbol include/adell_b200.h declares; the
ctypes mirror of adell_item matches the header's layout; host-only entry points behave."""

import ctypes as C
import os
import re

import numpy as np

from adell_mri_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "adell_b200.h")).read()
    return sorted(set(re.findall(r"\b(adell_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_item_layout_and_status_strings():
    lib = _lib.load()
    assert lib.adell_abi_version() == _lib.ABI_VERSION == 7
    assert lib.adell_item_size() == C.sizeof(_lib.Item) == 768
    assert b"no CUDA device" in lib.adell_status_string(-5)
    assert lib.adell_status_string(0) == b"ok"


def test_mat4_chain_matches_torch_cpu_products():
    import torch

    R = np.random.RandomState(0)
    mats = R.randn(7, 4, 4, 4).astype(np.float32)
    out = np.empty((7, 4, 4), np.float32)
    assert _lib.load().adell_mat4_chain(mats.ctypes.data, 7, 4, out.ctypes.data) == 0
    for b in range(7):
        ref = torch.from_numpy(mats[b, 0])
        for j in range(1, 4):
            ref = ref @ torch.from_numpy(mats[b, j])
        assert np.array_equal(ref.numpy(), out[b])


def test_prepare_validates_and_never_falls_back():
    lib = _lib.load()
    from adell_mri_b200.engine import aligned_items
    items = aligned_items(2)
    tiles = np.full(7, -1, np.int32)   # tile prefix (n + 1) + the launch's four chunk-queue words
    info = _lib.LaunchInfo()
    assert lib.adell_aug_prepare(items.ctypes.data, 2, tiles.ctypes.data, C.byref(info)) == -1  # zero shapes
    # identity items need no TMA descriptor, hence no driver: preparation succeeds on a CPU-only box
    buf = np.zeros(64, np.float32)
    for it in items:
        it["src"] = buf.ctypes.data
        it["dst"] = buf.ctypes.data
        it["src_shape"] = it["out_shape"] = it["grid_shape"] = (17, 33, 40)
        it["grid_sign"] = (1, 1, 1)
        it["flags"] = _lib.F_IDENTITY
    assert lib.adell_aug_prepare(items.ctypes.data, 2, tiles.ctypes.data, C.byref(info)) == 0
    assert info.total_tiles == 2 * 2 * 3 * 3 and list(tiles) == [0, 18, 36, 0, 0, 0, 0] and info.n_staged == 0


def test_misaligned_item_buffers_are_refused():
    """``adell_item`` is declared 64-byte aligned (include/adell_b200.h:107): the host entry points answer
    ADELL_ERR_ALIGN for a buffer that is not, instead of reading it with the aligned moves they were compiled with."""
    from adell_mri_b200.engine import ISZ, aligned_bytes
    lib = _lib.load()
    raw = aligned_bytes(2 * ISZ + 64)
    off = raw[16: 16 + 2 * ISZ]
    tiles = np.zeros(7, np.int32)
    info = _lib.LaunchInfo()
    assert lib.adell_aug_prepare(off.ctypes.data, 2, tiles.ctypes.data, C.byref(info)) == -3
    assert lib.adell_aug_plan(off.ctypes.data, 2, tiles.ctypes.data, C.byref(info)) == -3
    chains = np.zeros(2 * lib.adell_chain_size(), np.uint8)
    assert lib.adell_chain_compose(chains.ctypes.data, 2, off.ctypes.data) == -3
    assert lib.adell_aug_prepare(raw.ctypes.data, 0, tiles.ctypes.data, C.byref(info)) == 0   # empty launch: nothing is read
