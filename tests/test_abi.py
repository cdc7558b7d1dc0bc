"""CPU checks of the drop-in boundary: the shared object loads and exports exactly the entry points
include/fava_b200.h declares, and the ctypes table in fava_b200/_lib.py covers them all."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "fava_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fava_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("fava_init", "fava_plane_moments", "fava_moments_finalize", "fava_prolong", "fava_ke_spectrum",
                 "fava_stage_h2d", "fava_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in fava_b200.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from fava_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.fava_abi_version() == _lib.ABI_VERSION == 4
    assert lib.fava_launch_count() == 0


def test_struct_layouts_match_header():
    from fava_b200 import _lib

    assert ctypes.sizeof(_lib.LeafDesc) == 32
    assert ctypes.sizeof(_lib.ProlongLeaf) == 24


def test_no_device_fails_loudly(built_lib):
    """Without a GPU fava_init must fail with a message, never fall back to a CPU path."""
    import torch

    if torch.cuda.is_available():
        return
    from fava_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.fava_init(0, ctypes.byref(h))
    assert rc < 0
    assert lib.fava_last_error()
    import pytest
    from fava_b200 import device

    with pytest.raises(RuntimeError):
        device.get_context(0)
