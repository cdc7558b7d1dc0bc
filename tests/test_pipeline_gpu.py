"""GPU: the batch pipeline `python -m fava_b200` (reference fava/__main__.py) end to end on a synthetic run:
plt files -> Reynolds stresses + flame window -> window trajectory -> from_amr windows -> KE spectra, with the
result files, the checkpoint and a restart."""
import json

import numpy as np
import pytest

from fava_b200 import h5lite, synth

pytestmark = pytest.mark.gpu

L = 1.0e5  # cm per fine cell


def make_run(tmp_path, nfiles=3):
    bounds = ((-48 * L, 48 * L), (-16 * L, 16 * L), (-16 * L, 16 * L))
    mesh = synth.octree_mesh((6, 2, 2), (8, 8, 8), 2, seed=4, p_refine=0.6, bounds=bounds)
    assert mesh.fine_dims_xyz() == (96, 32, 32)
    rng = np.random.default_rng(0)
    for i in range(nfiles):
        xc = (10.0 + i) * L
        fields = {k: np.zeros((mesh.nblocks, 8, 8, 8), dtype=np.float32) for k in ("dens", "velx", "vely", "velz", "flam", "pres")}
        bb = mesh.bbox(np.float64)
        for b in range(mesh.nblocks):
            x = np.linspace(bb[b, 0, 0], bb[b, 0, 1], 17)[1::2][None, None, :]  # cell centres along x
            amp = np.exp(-(((x - xc) / (6 * L)) ** 2))
            fields["dens"][b] = 1.0 + 0.2 * rng.random((8, 8, 8))
            fields["velx"][b] = 0.1 * rng.standard_normal((8, 8, 8))
            fields["vely"][b] = amp * rng.standard_normal((8, 8, 8))
            fields["velz"][b] = amp * rng.standard_normal((8, 8, 8))
            fields["flam"][b] = 0.5 * (1.0 + np.tanh((x - xc) / L)) * np.ones((8, 8, 8))
            fields["pres"][b] = 1.0
        synth.write_flash_file(tmp_path / f"rt_hdf5_plt_cnt_{i:04d}", mesh, fields, time=0.1 * i)
    settings = {"data folder": str(tmp_path), "output folder": str(tmp_path), "basename": "rt_hdf5_plt_cnt", "dimension": 3,
                "model": "rt", "reynolds stress": {"skip": False}, "extract windows": {"skip": False},
                "fractal dimension": {"skip": False, "settings": {"field": "flam", "contours": 0.5}},
                "structure functions": {"skip": False, "settings": {"num_seps": 6, "num_points": 500, "sep_bounds": [1.0 * L, 12.0 * L],
                                                                     "log_scale": True, "anistropic": False}},
                "kinetic energy spectra": {"skip": False}}
    (tmp_path / "pipeline_settings.json").write_text(json.dumps(settings))
    return mesh


def test_pipeline_end_to_end_and_restart(cuda_device, tmp_path, capsys):
    import fava_b200 as fava
    from fava_b200.__main__ import main

    make_run(tmp_path)
    assert main(tmp_path) == 0
    out = capsys.readouterr().out
    assert "DONE!" in out
    ck = json.loads((tmp_path / "fava.checkpoint").read_text()) if (tmp_path / "fava.checkpoint").exists() else None
    assert ck is not None and ck["reynolds stress"] == {"index": 3} and ck["extract windows"] == {"index": 3}
    assert ck["analyze uniform data"]["index"] == 3 and ck["analyze uniform data"]["analysis"] is None

    model = fava.flash(tmp_path)
    assert model.nfiles(file_type="plt") == 3 and model.nfiles(file_type="uni") == 3 and model.nfiles(file_type="anl") == 3  # plt and uni results share one analysis file per number
    # stage 1 results equal a direct call; the window is 32 fine cells wide
    model.load(file_index=1, file_type="plt")
    radius, stress, means = model.reynolds_stress()
    with h5lite.File(tmp_path / "rt_hdf5_analysis_0001") as f:
        assert np.array_equal(f["reynolds stresses"]["tensor"]["Ryy"][()], stress["Ryy"])
        assert np.array_equal(f["reynolds stresses"]["radius"][()], radius)
        left, right = f["scalars"]["window left"][()], f["scalars"]["window right"][()]
        assert abs((right[0] - left[0]) - 32 * L) < 1e-6 * L and f["scalars"]["window dimensions"][()].tolist() == [32, 32, 32]
        assert float(f["scalars"]["time"][()]) == pytest.approx(0.1)
    # stage 3/4: 32^3 uniform windows and their spectra, equal to a direct call on the written file
    model.load(file_index=1, file_type="uni")
    assert (int(model.mesh.nxb), int(model.mesh.nyb), int(model.mesh.nzb)) == (32, 32, 32)
    sp = model.kinetic_energy_spectra()
    with h5lite.File(tmp_path / "rt_hdf5_analysis_0001") as f:
        for k in ("k", "total", "longitudinal", "transverse"):
            assert np.array_equal(f["kinetic energy spectra"][k][()], sp[k]), k
    fd = model.fractal_dimension("flam", 0.5)
    with h5lite.File(tmp_path / "rt_hdf5_analysis_0001") as f:
        for k, v in fd["flam"]["0.5"].items():
            assert np.array_equal(f["fractal dimension"]["flam"]["0.5"][k][()], v, equal_nan=True), k
        assert f["structure functions"]["longitudinal"]["3"][()].shape == (6,)
        assert np.all(f["structure functions"]["transverse"]["2"][()] > 0)
    # restart: everything is marked done, nothing is recomputed, cached results are re-used
    before = {p.name: p.stat().st_mtime_ns for p in tmp_path.glob("*uniform*")}
    assert main(tmp_path) == 0
    assert {p.name: p.stat().st_mtime_ns for p in tmp_path.glob("*uniform*")} == before
