"""CPU: the pure-Python HDF5 subset (fava_b200/h5lite.py) — round trips of every datatype the FLASH schema
uses, the h5py-shaped calls the reference makes, byte extents for staging."""
import numpy as np
import pytest

from fava_b200 import h5lite, synth
from fava_b200.util import HID_T


def test_roundtrip_all_flash_types(tmp_path):
    p = tmp_path / "t_hdf5_plt_cnt_0000"
    rng = np.random.default_rng(0)
    f32 = rng.random((3, 4, 5, 6)).astype(np.float32)
    f64 = rng.random((70, 70, 70))
    ints = np.arange(-5, 10, dtype=np.int32).reshape(3, 5)
    with h5lite.File(p, "w") as f:
        f.create_dataset("dens", data=f32)
        f.create_dataset(name="big ", shape=f64.shape, dtype="<f8", data=f64)
        f.create_dataset("gid", data=ints)
        f.create_dataset("unknown names", data=np.array([[b"dens"], [b"velx"]], dtype="S4"))
        f.create_dataset(name="integer scalars", shape=2, dtype=HID_T.I32_PARAMETER, data=[(f"{'nxb':256s}", 8), (f"{'nyb':256s}", 4)])
        f.create_dataset(name="real scalars", shape=1, dtype=HID_T.F64_PARAMETER, data=[(f"{'time':256s}", 0.25)])
        f.create_dataset(name="logical scalars", shape=1, dtype=HID_T.BOOL_PARAMETER, data=[(f"{'flag':256s}", 1)])
        f.create_dataset(name="string scalars", shape=1, dtype=HID_T.STR_PARAMETER, data=[(f"{'geometry':256s}", f"{'cartesian':256s}")])
        f.create_dataset(name="empty", shape=0, dtype=HID_T.I32_PARAMETER, data=[])
        for i in range(40):  # more objects than one symbol-table node holds
            f.create_dataset(f"extra{i:02d}", data=np.full(i + 1, i, dtype=np.int64))
        with pytest.raises(ValueError):
            f.create_dataset("dens", data=f32)
    assert h5lite.is_hdf5(p)
    with h5lite.File(p, "r") as f:
        assert len(list(f.keys())) == 49 and "dens" in f and "nope" not in f
        assert np.array_equal(f["dens"][()], f32) and f["dens"].dtype == np.float32 and f["dens"].shape == f32.shape
        assert np.array_equal(f["big "][()], f64) and f["big "].nbytes == f64.nbytes
        assert np.array_equal(f["gid"][()], ints)
        assert f["unknown names"][()].astype(str).tolist() == [["dens"], ["velx"]]
        t = f["integer scalars"]
        assert np.char.strip(t[:, "name"].astype(str)).tolist() == ["nxb", "nyb"] and t[:, "value"].tolist() == [8, 4]
        assert f["real scalars"][:, "value"][0] == 0.25
        b = f["logical scalars"]  # members stored out of order (value at 0, name at 4)
        assert b.dtype.itemsize == 260 and b[:, "value"][0] == 1 and b[:, "name"][0].strip() == b"flag"
        s = f["string scalars"]
        assert s[:, "value"][0].strip() == b"cartesian" and s.dtype.itemsize == 512
        assert f["empty"].shape == (0,)
        for i in (0, 17, 39):
            assert np.array_equal(f[f"extra{i:02d}"][()], np.full(i + 1, i))
        out = np.empty(ints.shape, dtype=np.int32)
        f["gid"].read_direct(out)
        assert np.array_equal(out, ints)
        off, nbytes = f["big "].extent()
        assert off % h5lite.DATA_ALIGN == 0 and nbytes == f64.nbytes
        with open(p, "rb") as raw:  # the extent really is the raw little-endian payload
            raw.seek(off)
            assert np.array_equal(np.frombuffer(raw.read(nbytes), dtype="<f8").reshape(f64.shape), f64)
        with pytest.raises(KeyError):
            f["nope"]


def test_rejects_non_hdf5_and_unsupported(tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"not an hdf5 file" * 10)
    assert not h5lite.is_hdf5(p)
    with pytest.raises(h5lite.H5LiteError):
        h5lite.File(p, "r")
    with pytest.raises(h5lite.H5LiteError):
        h5lite.File(tmp_path / "x", "q")
    with pytest.raises(FileNotFoundError):
        h5lite.File(tmp_path / "absent", "r+")
    with h5lite.File(tmp_path / "y", "w") as f:
        with pytest.raises(h5lite.H5LiteError):
            f.create_dataset("o", data=np.array([object()], dtype=object))
    with h5lite.File(tmp_path / "y", "r") as f:
        with pytest.raises(h5lite.H5LiteError):
            f.create_dataset("z", data=np.zeros(3))


def test_nested_groups_append_and_delete(tmp_path):
    """The analysis-result files of the reference (fava/model/model.py:138-185): nested groups, scalar datasets,
    append mode, overwrite by delete + create."""
    p = tmp_path / "res_hdf5_analysis_0000"
    with h5lite.File(p, "w") as f:
        g = f.create_group("reynolds stresses")
        t = g.create_group("tensor")
        t.create_dataset("Rxx", data=np.arange(5.0))
        t.create_dataset("Rxy", data=np.arange(5.0) * 2)
        g.create_dataset("radius", data=np.linspace(0, 1, 6))
        with pytest.raises(ValueError):
            f.create_group("reynolds stresses")  # the reference catches this and re-opens the group
        assert "tensor" in f["reynolds stresses"] and list(f["reynolds stresses"]["tensor"].keys()) == ["Rxx", "Rxy"]
    with h5lite.File(p, "a") as f:  # append: scalars group, overwrite one dataset
        s = f.create_group("scalars")
        s.create_dataset("time", data=0.125)
        s.create_dataset("window dimensions", data=np.array([3, 4, 5]))
        t = f["reynolds stresses"]["tensor"]
        del t["Rxy"]
        t.create_dataset("Rxy", data=np.ones(5))
        with pytest.raises(KeyError):
            del t["nope"]
    with h5lite.File(p, "r") as f:
        assert sorted(f.keys()) == ["reynolds stresses", "scalars"]
        assert np.array_equal(f["reynolds stresses"]["tensor"]["Rxx"][()], np.arange(5.0))
        assert np.array_equal(f["reynolds stresses/tensor/Rxy"][()], np.ones(5))
        assert np.array_equal(f["reynolds stresses"]["radius"][()], np.linspace(0, 1, 6))
        tm = f["scalars"]["time"]
        assert tm.shape == () and float(tm[()]) == 0.125
        assert f["scalars"]["window dimensions"][()].tolist() == [3, 4, 5]
    # 20 sub-groups with 12 datasets each: several symbol-table nodes per group, recursion
    q = tmp_path / "many"
    with h5lite.File(q, "w") as f:
        for i in range(20):
            g = f.create_group(f"g{i:02d}")
            for j in range(12):
                g.create_dataset(f"d{j:02d}", data=np.full(j + 1, i * 100 + j, dtype=np.int32))
    with h5lite.File(q) as f:
        assert len(list(f.keys())) == 20
        assert f["g13"]["d07"][()].tolist() == [1307] * 8


def test_synthetic_flash_file_schema(tmp_path):
    """The generator writes exactly what the reference's loader reads."""
    mesh = synth.octree_mesh((2, 1, 1), (4, 4, 4), 3, seed=1)
    fields = synth.block_fields(mesh, names=("dens", "velx"))
    p = tmp_path / "s_hdf5_plt_cnt_0000"
    synth.write_flash_file(p, mesh, fields)
    with h5lite.File(p) as f:
        nb = mesh.nblocks
        assert f["dens"].shape == (nb, 4, 4, 4) and f["dens"].dtype == np.float32
        assert f["bounding box"].shape == (nb, 3, 2) and f["bounding box"].dtype == np.float32
        assert f["gid"].shape == (nb, 15) and f["refine level"][()].max() == 3
        nt, gid = f["node type"][()], f["gid"][()]
        assert np.all((nt == 1) == (gid[:, 7] < 0))  # leaves have no children
        lev = f["refine level"][()]
        kids = gid[nt != 1, 7:15] - 1
        assert np.all(lev[kids] == lev[nt != 1][:, None] + 1)
        # parents hold the restriction of their children
        b = int(np.flatnonzero(nt != 1)[0])
        d = f["dens"][()].astype(np.float64)
        child0 = d[gid[b, 7] - 1].reshape(2, 2, 2, 2, 2, 2).mean(axis=(1, 3, 5))
        assert np.allclose(d[b][:2, :2, :2], child0, rtol=1e-6)


def test_append_survives_a_failed_flush_and_an_unreadable_file(tmp_path, monkeypatch):
    """The result cache (Model.save_to_hdf5, mode "a") re-writes the file on close: a failure in the middle must leave
    the previous contents intact (write-to-temp + rename), and a corrupt file must not block later runs."""
    import os

    path = tmp_path / "results.h5"
    with h5lite.File(path, "w") as f:
        f.create_dataset("a", data=np.arange(5.0))
    real_replace = os.replace

    def boom(src, dst):
        raise OSError("disk full")

    monkeypatch.setattr(os, "replace", boom)
    try:
        with h5lite.File(path, "a") as f:
            f.create_dataset("b", data=np.arange(3.0))
    except OSError:
        pass
    monkeypatch.setattr(os, "replace", real_replace)
    assert [p.name for p in tmp_path.iterdir()] == ["results.h5"]  # no temp file left behind
    with h5lite.File(path) as f:
        assert np.array_equal(f["a"][()], np.arange(5.0)) and "b" not in f
    path.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 40)  # truncated
    with h5lite.File(path, "a") as f:
        f.create_dataset("c", data=np.arange(2.0))
    with h5lite.File(path) as f:
        assert np.array_equal(f["c"][()], np.arange(2.0))
