"""Host-side leaf / prolongation tables (no GPU): the arrays handed to the C ABI are private, read-only copies, and
every table object carries its own uid - the key of the library's device-table cache
(fava_plane_moments_blocks_uid, include/fava_b200.h)."""
import ctypes as C

import numpy as np
import pytest

from fava_b200 import _lib, device


def test_leaf_table_is_an_immutable_copy_with_a_fresh_uid():
    blocks = np.arange(5)
    ilo = np.array([0, 8, 16, 8, 0])
    scale = np.array([1, 1, 2, 1, 4])
    vf = np.linspace(0.1, 0.5, 5)
    a = device.leaf_table(blocks, ilo, scale, vf)
    b = device.leaf_table(blocks, ilo, scale, vf)
    assert a.uid != b.uid and a.uid > 0 and b.uid > 0
    assert a.n == 5 and a.arr.dtype == device.LEAF_DTYPE and a.arr.dtype.itemsize == C.sizeof(_lib.LeafDesc) == 32
    assert not a.arr.flags.writeable
    with pytest.raises(ValueError):
        a.arr["ilo"][0] = 3
    ilo[0] = 99  # the caller's arrays are not aliased
    assert int(a.arr["ilo"][0]) == 0
    # the ctypes view reads the same bytes
    assert [a.ptr[i].scale for i in range(5)] == [1, 1, 2, 1, 4] and a.ptr[4].vol_frac == pytest.approx(0.5)


def test_empty_tables_have_a_null_pointer():
    t = device.leaf_table([], [], [], [])
    assert t.n == 0 and not bool(t.ptr)
