"""Small host utilities mirroring fava/util: the `timer` decorator (util/__init__.py:7-16), the on-disk
dtype table the FLASH writer uses (util/_types.py:5-26) and the exception types (util/_exceptions.py)."""

from __future__ import annotations

import time

from fava_b200 import dist


def timer(func):
    """Print `Timing: <name> --> <seconds>` on the root rank, like the reference's decorator."""

    def timed(*args, **kwargs):
        t0 = time.perf_counter()
        result = func(*args, **kwargs)
        dt = time.perf_counter() - t0
        if dist.is_root():
            print(f"Timing: {func.__name__} --> {dt:2.4f}", flush=True)
        return result

    timed.__name__ = getattr(func, "__name__", "timed")
    timed.__doc__ = getattr(func, "__doc__", None)
    timed.__wrapped__ = func
    return timed


class HID_T:
    """HDF5 element types of a FAVA-written FLASH file."""

    F32 = "<f4"
    F64 = "<f8"
    I32 = "<i4"
    I64 = "<i8"
    F64_PARAMETER = [("name", "S256"), ("value", "<f8")]
    I32_PARAMETER = [("name", "S256"), ("value", "<i4")]
    BOOL_PARAMETER = {"names": ["name", "value"], "formats": ["S256", "<i4"], "offsets": [4, 0], "itemsize": 260}
    STR_PARAMETER = {"names": ["name", "value"], "formats": ["S256", "S256"], "offsets": [256, 0], "itemsize": 512}
    UNKNOWN_NAMES = "S4"


class NotCallableError(TypeError):
    def __init__(self, obj):
        super().__init__(f"{obj!r} is not callable and cannot be registered as an analysis")


class MeshNotLoadedError(RuntimeError):
    pass
