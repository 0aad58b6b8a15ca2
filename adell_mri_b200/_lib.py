"""ctypes binding of the C-ABI shared library (include/adell_b200.h).

The library is built in-tree by ``adell_mri_b200/csrc/build.sh`` (nvcc, sm_100a only).
There is deliberately NO fallback: if the library is missing or a call returns a
non-zero status a ``RuntimeError`` is raised.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADELL_B200_LIB") or os.path.join(_HERE, "libadell_b200.so")  # override: tooling only

# constants mirrored from include/adell_b200.h
F32, I16, U8 = 0, 1, 2
NEAREST, TRILINEAR = 0, 1
PAD_ZEROS, PAD_BORDER, PAD_REFLECTION = 0, 1, 2
F_IDENTITY, F_CLIP, F_STRICT, F_PHILOX, F_PRE_DEV, F_TMAP, F_FASTCOORD, F_WIN_DEV = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80
SCALER_MINMAX, SCALER_ADC_SEG, SCALER_ADC_CLASS, SCALER_RANGE, SCALER_ZSCORE = 0, 1, 2, 3, 4

PADDING_MODES = {"zeros": PAD_ZEROS, "border": PAD_BORDER, "reflection": PAD_REFLECTION}
INTERP_MODES = {"nearest": NEAREST, "bilinear": TRILINEAR, "trilinear": TRILINEAR, "linear": TRILINEAR}


class Item(C.Structure):
    """Mirror of ``adell_item`` (768 bytes, 64-byte aligned)."""

    _fields_ = [
        ("tmap", C.c_uint8 * 128),
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("noise", C.c_void_p),
        ("pre_dev", C.c_void_p),
        ("win_dev", C.c_void_p),
        ("src_stride", C.c_int64 * 3),
        ("dst_stride", C.c_int64 * 3),
        ("src_shape", C.c_int32 * 3),
        ("src_vlo", C.c_int32 * 3),
        ("src_vhi", C.c_int32 * 3),
        ("out_shape", C.c_int32 * 3),
        ("grid_shape", C.c_int32 * 3),
        ("grid_off", C.c_int32 * 3),
        ("grid_sign", C.c_int32 * 3),
        ("out_vlo", C.c_int32 * 3),
        ("out_vhi", C.c_int32 * 3),
        ("tmap_off", C.c_int32 * 3),
        ("tmap_sign", C.c_int32 * 3),
        ("tmap_box", C.c_int32 * 3),
        ("A", C.c_float * 12),
        ("nrm", C.c_float * 3),
        ("pre_scale", C.c_float),
        ("pre_offset", C.c_float),
        ("clip_lo", C.c_float),
        ("clip_hi", C.c_float),
        ("post_scale", C.c_float),
        ("post_offset", C.c_float),
        ("noise_std", C.c_float),
        ("philox_seed", C.c_uint64),
        ("philox_offset", C.c_uint64),
        ("src_dtype", C.c_uint8),
        ("interp", C.c_uint8),
        ("padding", C.c_uint8),
        ("flags", C.c_uint8),
        # derived by adell_aug_prepare (callers leave them zero)
        ("tile_dim", C.c_uint8 * 3),
        ("kind", C.c_uint8),
        ("n_tiles", C.c_int32 * 3),
        ("fp_smin", C.c_float * 3),
        ("fp_smax", C.c_float * 3),
        ("fp_fix", C.c_int32),
        ("fp_U0", C.c_double * 3),
        ("fp_D", C.c_double * 9),
        ("shear", C.c_int8 * 32),
        ("dmap", C.c_uint8 * 128),
    ]


ITEM_SIZE = C.sizeof(Item)
assert ITEM_SIZE == 768, ITEM_SIZE


class Chain(C.Structure):
    """Mirror of ``adell_chain`` (host-side chain descriptor consumed by ``adell_chain_prepare_steps``)."""

    _fields_ = [
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("pre_dev", C.c_void_p),
        ("win_dev", C.c_void_p),
        ("src_stride", C.c_int64 * 3),
        ("dst_stride", C.c_int64 * 3),
        ("src_shape", C.c_int32 * 3),
        ("crop0_start", C.c_int32 * 3),
        ("crop0_size", C.c_int32 * 3),
        ("crop1_size", C.c_int32 * 3),
        ("A", C.c_float * 12),
        ("pre_scale", C.c_float),
        ("pre_offset", C.c_float),
        ("post_scale", C.c_float),
        ("post_offset", C.c_float),
        ("src_dtype", C.c_uint8),
        ("interp", C.c_uint8),
        ("padding", C.c_uint8),
        ("flags", C.c_uint8),
        ("flip0", C.c_uint8),
        ("flip1", C.c_uint8),
        ("reserved_", C.c_uint8 * 2),
    ]


CHAIN_AFFINE, CHAIN_STRICT = 0x01, 0x02


class SeqOp(C.Structure):
    """Mirror of ``adell_seq_op`` (one member of a volume's ordered op list)."""

    _fields_ = [
        ("A", C.c_float * 12),
        ("scale", C.c_double),
        ("offset", C.c_double),
        ("philox_seed", C.c_uint64),
        ("philox_offset", C.c_uint64),
        ("philox_std", C.c_float),
        ("kind", C.c_uint8),
        ("interp", C.c_uint8),
        ("padding", C.c_uint8),
        ("reserved_", C.c_uint8),
    ]


SEQ_MAX_OPS = 4
OP_NONE, OP_AFFINE, OP_INTENSITY, OP_PHILOX = 0, 1, 2, 3
SEQ_FAST, SEQ_STRICT = 0x01, 0x02
ERR_NO_SPACE = -8


class Seq(C.Structure):
    """Mirror of ``adell_seq`` (host-side sequence descriptor consumed by ``adell_seq_prepare_steps``)."""

    _fields_ = [
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("src_stride", C.c_int64 * 3),
        ("dst_stride", C.c_int64 * 3),
        ("src_shape", C.c_int32 * 3),
        ("crop0_start", C.c_int32 * 3),
        ("crop0_size", C.c_int32 * 3),
        ("src_dtype", C.c_uint8),
        ("n_ops", C.c_uint8),
        ("flags", C.c_uint8),
        ("reserved_", C.c_uint8),
        ("ops", SeqOp * SEQ_MAX_OPS),
    ]


class PosNeg(C.Structure):
    """Mirror of ``adell_posneg``."""

    _fields_ = [("indices", C.c_void_p), ("pick", C.c_int64), ("shape", C.c_int32 * 3), ("size", C.c_int32 * 3)]


assert C.sizeof(PosNeg) == 40


class LaunchInfo(C.Structure):
    """Mirror of ``adell_launch_info``."""

    _fields_ = [("total_tiles", C.c_int64), ("smem_bytes", C.c_int32), ("n_staged", C.c_int32), ("first_copy_tile", C.c_int64)]


class SeqLaunch(C.Structure):
    """Mirror of ``adell_seq_launch``."""

    _fields_ = [("item_off", C.c_int64), ("n_items", C.c_int32), ("step", C.c_int32), ("info", LaunchInfo)]


class Vol(C.Structure):
    """Mirror of ``adell_vol``."""

    _fields_ = [("data", C.c_void_p), ("n", C.c_int64), ("dtype", C.c_int32), ("_pad", C.c_int32)]


assert C.sizeof(Vol) == 24

_SIGNATURES = {
    "adell_abi_version": (C.c_int, []),
    "adell_status_string": (C.c_char_p, [C.c_int]),
    "adell_item_size": (C.c_int, []),
    "adell_device_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "adell_mat4_chain": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "adell_affine_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "adell_aug_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adell_aug_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adell_aug_prepare_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adell_aug_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adell_chain_size": (C.c_int, []),
    "adell_chain_compose": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "adell_chain_prepare_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "adell_seq_size": (C.c_int, []),
    "adell_seq_prepare_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "adell_aug_gather_launches": (C.c_int, []),
    "adell_minmax": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "adell_meanstd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adell_intensity_map": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p],
    ),
    "adell_label_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "adell_mask_bbox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "adell_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "adell_rician_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "adell_gamma_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "adell_scaler_coefs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "adell_coefs_to_affine": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adell_hist_pass": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "adell_hist_select": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adell_quantile_workspace": (C.c_int64, [C.c_int, C.c_int, C.c_int64, C.c_int64]),
    "adell_quantile_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "adell_quantile_fell_back": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    "adell_posneg_starts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adell_mixup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "adell_percentile_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

ABI_VERSION = 7
_lib = None


def _require_fma() -> None:
    """The host composers are compiled with ``-mfma`` (csrc/build.sh): refuse a host CPU without the instruction
    instead of dying on it."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        return
    if flags and " fma" not in flags:
        raise RuntimeError("libadell_b200.so needs a host CPU with FMA (x86-64-v3); rebuild csrc/build.sh without -mfma")


def load() -> C.CDLL:
    """Load ``libadell_b200.so`` (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with adell_mri_b200/csrc/build.sh "
            "(or __graft_entry__.build()); there is no CPU fallback."
        )
    _require_fma()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.adell_abi_version() != ABI_VERSION:
        raise RuntimeError("adell_b200 ABI version mismatch")
    if lib.adell_item_size() != C.sizeof(Item):
        raise RuntimeError("adell_item layout mismatch between header and binding")
    if lib.adell_chain_size() != C.sizeof(Chain):
        raise RuntimeError("adell_chain layout mismatch between header and binding")
    if lib.adell_seq_size() != C.sizeof(Seq):
        raise RuntimeError("adell_seq layout mismatch between header and binding")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().adell_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")
