"""CPU: the host mirror of the reference's model layer (fava/model/model.py, fava/model/flash.py) — registries,
file discovery, filename conversion, the analysis-result writer.  No GPU is touched."""
import numpy as np
import pytest

import fava_b200 as fava
from fava_b200 import h5lite, synth


def _touch(d, name):
    (d / name).write_bytes(b"x")


def test_registries_and_entry_points():
    assert {"FLASH", "FlashUniform", "Structured", "Unstructured"} <= set(fava.Model.mesh_names())
    for name in ("reynolds_stress", "favre_stress", "kinetic_energy_spectra", "slice_average", "slice_integral", "from_amr"):
        assert callable(getattr(fava.Model, name))
    assert callable(fava.flash) and fava.FLASH.__name__ == "FLASH"

    @fava.Model.register_analysis()
    def reynolds_stress(self):  # an existing name is NOT replaced unless overwrite=True (reference model.py:124)
        return "shadow"

    assert fava.Model.reynolds_stress.__name__ == "reynolds_stress" and fava.Model.reynolds_stress is not reynolds_stress
    with pytest.raises(TypeError):
        fava.Model.register_analysis()(42)
    assert fava.mesh.Mesh().mesh_type == "Mesh" and fava.mesh.Mesh.is_this_your_mesh() is False
    assert fava.mesh.FLASH.is_this_your_mesh("run_hdf5_plt_cnt_0001") and fava.FlashUniform.is_this_your_mesh("a_hdf5_uniform_0003")


def test_flash_model_file_discovery(tmp_path):
    with pytest.raises(FileNotFoundError):
        fava.flash(tmp_path / "missing")
    with pytest.raises(FileNotFoundError):
        fava.flash(tmp_path)  # empty directory
    for n in ("r_hdf5_plt_cnt_0000", "r_hdf5_plt_cnt_0010", "r_hdf5_chk_0002", "r_hdf5_uniform_0010", "r_hdf5_part_0001", "notes.txt"):
        _touch(tmp_path, n)
    m = fava.flash(tmp_path, name="run")
    assert m.name == "run" and fava.flash(tmp_path).name == tmp_path.name
    assert m.nfiles(file_type="plt") == 2 and m.nfiles(file_type="chk") == 1 and m.nfiles(file_type="uni") == 1
    assert sorted(m.plt_files["by number"]) == [0, 10] and m.plt_files["by index"][1].name == "r_hdf5_plt_cnt_0010"
    assert m.convert_filename_type("plt", "uni") is None  # nothing loaded yet
    with pytest.raises(AssertionError):
        m.load(file_number=5, file_type="plt")
    with pytest.raises(NotImplementedError):
        m.load(file_index=0, file_type="prt")


def test_mesh_metadata_load_without_gpu(tmp_path):
    """`load()` only reads metadata (h5lite); field staging is what needs the device."""
    mesh = synth.octree_mesh((2, 1, 1), (4, 4, 4), 2, seed=2)
    fields = synth.block_fields(mesh, names=("dens", "velx"))
    synth.write_flash_file(tmp_path / "m_hdf5_plt_cnt_0003", mesh, fields, time=1.5)
    model = fava.flash(tmp_path)
    model.load(file_number=3, file_type="plt")
    m = model.mesh
    assert (int(m.ndim), int(m.nxb), int(m.nblocks)) == (3, 4, mesh.nblocks) and float(m.time) == 1.5
    assert m.fields.tolist() == ["dens", "velx"] and m.block_bounds.dtype == np.float32
    assert m.refine_level.dtype == np.int64 and int(m.refine_level_max) == 2
    assert np.array_equal(m.get_blocklist("LEAF"), np.flatnonzero(mesh.node_type == 1))
    assert model.convert_filename_type("plt", "uni").name == "m_hdf5_uniform_0003"
    assert m.data("no such field") is None
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):  # no CPU fallback: staging needs the device
            m.load_data(["dens"])


def test_save_to_hdf5_roundtrip(tmp_path):
    _touch(tmp_path, "r_hdf5_plt_cnt_0000")
    m = fava.Model(tmp_path)
    fn = tmp_path / "r_hdf5_analysis_0000"
    m.save_to_hdf5({"reynolds stresses": {"tensor": {"Rxx": np.arange(3.0)}, "radius": np.arange(4.0)}}, fn)
    m.save_to_hdf5({"scalars": {"time": 2.0}, "reynolds stresses": {"tensor": {"Rxx": np.zeros(3)}}}, fn)
    assert m.hdf5_key_exists("scalars", fn) and not m.hdf5_key_exists("x", fn) and not m.hdf5_key_exists("x", tmp_path / "none")
    with h5lite.File(fn) as f:
        assert np.array_equal(f["reynolds stresses"]["tensor"]["Rxx"][()], np.zeros(3))
        assert np.array_equal(f["reynolds stresses"]["radius"][()], np.arange(4.0))
        assert float(f["scalars"]["time"][()]) == 2.0
