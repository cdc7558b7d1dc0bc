#!/usr/bin/env python
"""Generate tests/golden/g8_pipeline.npz: the UNMODIFIED reference's batch pipeline (`python -m fava`,
fava/__main__.py) run end to end under the import shims of oracle/ref_harness.py on a small synthetic FLASH run
(three AMR plt files with a moving flame front).  The fixture holds the inputs (mesh + block fields + settings +
the np.random seed) and every dataset of the result files the reference wrote: the analysis files (stresses, flame
window, fractal dimension, structure functions, spectra), two fields of each uniform window file, and the checkpoint.

    python tests/golden/make_pipeline_golden.py
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from fava_b200 import h5lite, synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = Path(__file__).resolve().parent
L = 2.0e5  # cm per fine cell: the reference hard-codes a 32e5 cm window (16 cells here)
RANDOM_FIELDS = ("dens", "velx", "vely", "velz", "flam")
CONSTANT_FIELDS = ("pres", "temp", "divv", "igtm", "vort")  # the reference's window extraction assumes they exist
SEED = 5
NFILES = 3


def build_inputs():
    bounds = ((-48 * L, 48 * L), (-8 * L, 8 * L), (-8 * L, 8 * L))
    mesh = synth.octree_mesh((6, 1, 1), (8, 8, 8), 2, seed=4, p_refine=0.6, bounds=bounds)
    rng = np.random.default_rng(0)
    files = []
    for i in range(NFILES):
        xc = (4.0 + 2 * i) * L
        fields = {k: np.zeros((mesh.nblocks, 8, 8, 8), dtype=np.float32) for k in RANDOM_FIELDS}
        bb = mesh.bbox(np.float64)
        for b in range(mesh.nblocks):
            x = np.linspace(bb[b, 0, 0], bb[b, 0, 1], 17)[1::2][None, None, :]  # cell centres along x
            amp = np.exp(-(((x - xc) / (6 * L)) ** 2))
            fields["dens"][b] = 1.0 + 0.2 * rng.random((8, 8, 8))
            fields["velx"][b] = 0.1 * rng.standard_normal((8, 8, 8))
            fields["vely"][b] = amp * rng.standard_normal((8, 8, 8))
            fields["velz"][b] = amp * rng.standard_normal((8, 8, 8))
            fields["flam"][b] = 0.5 * (1.0 + np.tanh((x - xc) / L)) * np.ones((8, 8, 8))
        files.append(fields)
    settings = {"basename": "rt_hdf5_plt_cnt", "dimension": 3, "model": "rt", "reynolds stress": {"skip": False},
                "extract windows": {"skip": False},
                "fractal dimension": {"skip": False, "settings": {"field": "flam", "contours": 0.5}},
                "structure functions": {"skip": False, "settings": {"num_seps": 4, "num_points": 200,
                                                                     "sep_bounds": [1.0 * L, 6.0 * L], "log_scale": True,
                                                                     "anistropic": False}},
                "kinetic energy spectra": {"skip": False}}
    return mesh, files, settings


def write_run(tmp: Path, mesh, files, settings) -> None:
    """The run directory both drivers read (also used by tests/test_pipeline_golden_gpu.py)."""
    for i, fields in enumerate(files):
        full = dict(fields)
        for k in CONSTANT_FIELDS:
            full[k] = np.ones_like(fields["dens"])
        synth.write_flash_file(tmp / f"rt_hdf5_plt_cnt_{i:04d}", mesh, full, time=0.1 * i)
    s = dict(settings)
    s["data folder"] = s["output folder"] = str(tmp)
    (tmp / "pipeline_settings.json").write_text(json.dumps(s))


def flatten(group, prefix: str, out: dict) -> None:
    for k in group.keys():
        node = group[k]
        if hasattr(node, "keys"):
            flatten(node, f"{prefix}{k}/", out)
        else:
            out[f"{prefix}{k}"] = np.asarray(node[()])


def collect_results(tmp: Path) -> dict:
    out = {}
    for i in range(NFILES):
        with h5lite.File(tmp / f"rt_hdf5_analysis_{i:04d}") as f:
            flatten(f, f"anl{i}:", out)
        with h5lite.File(tmp / f"rt_hdf5_uniform_{i:04d}") as f:
            for k in ("dens", "flam", "bounding box"):
                out[f"uni{i}:{k}"] = np.asarray(f[k][()])
    return out


def main():
    if not rh.reference_available():
        raise SystemExit("the reference is not mounted; goldens can only be generated in the build container")
    mesh, files, settings = build_inputs()
    tmp = Path(tempfile.mkdtemp(prefix="fava_pipeline_golden_"))
    write_run(tmp, mesh, files, settings)
    rh.ref_modules()
    os.chdir(tmp)  # the reference binds ./pipeline_settings.json and ./fava.checkpoint at import time
    spec = importlib.util.spec_from_file_location("ref_fava_main", rh.REFERENCE_ROOT / "fava" / "__main__.py")
    ref_main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_main)
    np.random.seed(SEED)  # structure_functions draws from the global stream
    ref_main.main()
    out = collect_results(tmp)
    ck = json.loads((tmp / "fava.checkpoint").read_text())
    out["checkpoint_progress"] = np.array(json.dumps({k: v for k, v in ck.items() if k != "settings"}, sort_keys=True))
    out["settings"] = np.array(json.dumps(settings))
    out["seed"] = np.array(SEED)
    for k, v in dict(nb_xyz=np.array([mesh.nxb, mesh.nyb, mesh.nzb]), nroot=np.array(mesh.nroot), bounds=mesh.bounds,
                     level=mesh.level, origin=mesh.origin, node_type=mesh.node_type, gid=mesh.gid,
                     which_child=mesh.which_child).items():
        out[f"mesh_{k}"] = v
    for i, fields in enumerate(files):
        for k, v in fields.items():
            out[f"in{i}_{k}"] = v
    np.savez_compressed(OUT / "g8_pipeline", **out)
    print(f"wrote g8_pipeline.npz ({(OUT / 'g8_pipeline.npz').stat().st_size / 1024:.0f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
