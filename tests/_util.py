"""Shared helpers of the test-suite: golden fixtures, mesh reconstruction, the parity norm."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from fava_b200 import synth
from oracle import fava_oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"
FIELDS = ("dens", "velx", "vely", "velz")
STRESS = ("Rxx", "Rxy", "Rxz", "Ryy", "Ryz", "Rzz")
RTOL = 1e-12  # BASELINE.json north_star: fp64 profiles and spectra within 1e-12 relative (max-norm per array)


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def golden_mesh(g: dict) -> synth.SynthMesh:
    return synth.SynthMesh(g["mesh_nb_xyz"], g["mesh_nroot"], g["mesh_bounds"], g["mesh_level"], g["mesh_origin"],
                           g["mesh_node_type"], g["mesh_gid"], g["mesh_which_child"])


def golden_fields(g: dict, names=FIELDS) -> dict:
    return {k: g[f"in_{k}"] for k in names if f"in_{k}" in g}


def oracle_geom(mesh: synth.SynthMesh, bbox_dtype=np.float32) -> orc.MeshGeom:
    return orc.MeshGeom((mesh.nxb, mesh.nyb, mesh.nzb), mesh.nroot, mesh.bounds, mesh.bbox(bbox_dtype), mesh.level,
                        mesh.node_type)


def oracle_data(fields: dict) -> dict:
    """file layout [...,z,y,x] -> the reference's in-memory float64 [blk,i,j,k]."""
    out = {}
    for k, v in fields.items():
        a = orc.load_like_reference(v)
        out[k] = a if a.ndim == 4 else a[None, ...]
    return out


def maxnorm_close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    if a.size == 0:
        return
    scale = np.nanmax(np.abs(b)) if np.isfinite(b).any() else 0.0
    err = np.nanmax(np.abs(a - b)) if np.isfinite(b).any() else 0.0
    assert err <= rtol * scale + 1e-300, f"{what}: max|a-b|={err:.3e} > {rtol:g}*max|b|={scale:.3e}"
