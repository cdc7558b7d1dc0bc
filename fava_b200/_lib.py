"""ctypes binding of libfava_b200.so (the C ABI declared in include/fava_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails, a RuntimeError is raised
(the reference's convention is `logger.exception` + `raise RuntimeError`, _flash.py:161-163).
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfava_b200.so"

FAVA_F32 = 0
FAVA_F64 = 1
FAVA_NMOM = 14
FAVA_FRACTAL_MAXLEVELS = 32

c_void_p = C.c_void_p
c_int = C.c_int
c_i64 = C.c_int64
c_double = C.c_double
c_double_p = C.POINTER(C.c_double)


class LeafDesc(C.Structure):
    """struct fava_leaf_desc"""

    _fields_ = [
        ("block", C.c_int64),
        ("ilo", C.c_int64),
        ("scale", C.c_int32),
        ("pad_", C.c_int32),
        ("vol_frac", C.c_double),
    ]


class ProlongLeaf(C.Structure):
    """struct fava_prolong_leaf"""

    _fields_ = [
        ("block", C.c_int64),
        ("off", C.c_int32 * 3),
        ("scale", C.c_int32),
    ]


# name -> (restype, argtypes); exactly the entry points of include/fava_b200.h
SIGNATURES: dict[str, tuple] = {
    "fava_init": (c_int, [c_int, C.POINTER(c_void_p)]),
    "fava_shutdown": (c_int, [c_void_p]),
    "fava_stream_sync": (c_int, [c_void_p, c_void_p]),
    "fava_last_error": (C.c_char_p, []),
    "fava_abi_version": (c_int, []),
    "fava_launch_count": (c_i64, []),
    "fava_plane_pivots": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p],
    ),
    "fava_plane_moments": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p,
         c_int, c_void_p],
    ),
    "fava_plane_moments_xz": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p],
    ),
    "fava_plane_moments_xyz_supported": (c_int, [c_i64, c_i64, c_i64]),
    "fava_plane_moments_xyz": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "fava_plane_moments_blocks": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int,
         C.POINTER(LeafDesc), c_i64, c_i64, c_void_p, c_void_p, c_void_p],
    ),
    "fava_plane_moments_blocks_uid": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int,
         C.POINTER(LeafDesc), c_i64, C.c_uint64, c_i64, c_void_p, c_void_p, c_void_p],
    ),
    "fava_moments_repivot": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "fava_moments_finalize": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_i64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "fava_plane_sum": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p]),
    "fava_plane_sum_blocks": (
        c_int,
        [c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, C.POINTER(LeafDesc), c_i64, c_i64, c_void_p, c_void_p],
    ),
    "fava_plane_sum_blocks_uid": (
        c_int,
        [c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, C.POINTER(LeafDesc), c_i64, C.c_uint64, c_i64, c_void_p,
         c_void_p],
    ),
    "fava_prolong": (
        c_int,
        [c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, C.POINTER(ProlongLeaf), c_i64, c_i64, c_i64, c_i64,
         c_void_p, c_void_p],
    ),
    "fava_ke_spectrum": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_double_p, c_double_p, c_double_p,
         c_double_p, c_void_p],
    ),
    "fava_ke_weight3": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p,
         c_void_p],
    ),
    "fava_fft_native_supported": (c_int, [c_i64]),
    "fava_spectral_pitch": (c_i64, [c_i64]),
    "fava_ke_transform_x": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "fava_ke_transform_y": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p]),
    "fava_ke_transform_z": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p]),
    "fava_fft_x_weight3": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p,
         c_void_p],
    ),
    "fava_fft_cols": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_void_p, c_void_p]),
    "fava_fft_y_scatter": (
        c_int,
        [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p],
    ),
    "fava_a2a_pack": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p],
    ),
    "fava_spectrum_bin": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_double, c_void_p, c_void_p],
    ),
    "fava_spectrum_finalize": (
        c_int,
        [c_void_p, c_void_p, c_i64, c_double_p, c_double_p, c_double_p, c_double_p, c_void_p],
    ),
    "fava_fractal_tiles": (
        c_int,
        [c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_double, c_void_p, c_void_p,
         c_void_p],
    ),
    "fava_fractal_coarse": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p]),
    "fava_sf_gather": (
        c_int,
        [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_double_p,
         c_double_p, c_void_p, c_void_p, c_void_p],
    ),
    "fava_sf_moments": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_int, c_void_p, c_void_p],
    ),
    "fava_stage_h2d": (c_int, [c_void_p, C.c_char_p, c_i64, c_i64, c_void_p, c_void_p]),
    "fava_stage_host_h2d": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "fava_workspace": (c_int, [c_void_p, c_int, c_i64, C.POINTER(c_void_p)]),
    "fava_ipc_export": (c_int, [c_void_p, C.POINTER(C.c_ubyte * 64)]),
    "fava_ipc_open": (c_int, [C.POINTER(C.c_ubyte * 64), C.POINTER(c_void_p)]),
    "fava_ipc_close": (c_int, [c_void_p]),
}

ABI_VERSION = 4  # include/fava_b200.h: FAVA_ABI_VERSION
_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libfava_b200.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m fava_b200.build` "
            "(there is no CPU fallback for the FAVA hot path)"
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.fava_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libfava_b200 ABI version {lib.fava_abi_version()} != {ABI_VERSION}: rebuild with "
                           "`python -m fava_b200.build`")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Turn a negative status into RuntimeError carrying fava_last_error()."""
    if rc != 0:
        msg = load().fava_last_error()
        raise RuntimeError(f"libfava_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")
