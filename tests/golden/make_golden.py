#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under the import
shims of oracle/ref_harness.py on deterministic synthetic FLASH files.  Run in the build container:

    python tests/golden/make_golden.py

Each .npz holds the synthetic INPUT (block arrays + mesh metadata, so the fixture does not depend on
libm reproducing the generator bit for bit) and the reference's OUTPUT.  The reference ships no golden
vectors of its own for this path (its tests only touch class names), so these are the parity pins.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from fava_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = Path(__file__).resolve().parent
FIELDS = ("dens", "velx", "vely", "velz")


def mesh_arrays(mesh: synth.SynthMesh) -> dict:
    return dict(nb_xyz=np.array([mesh.nxb, mesh.nyb, mesh.nzb]), nroot=np.array(mesh.nroot), bounds=mesh.bounds,
                level=mesh.level, origin=mesh.origin, node_type=mesh.node_type, gid=mesh.gid,
                which_child=mesh.which_child)


def save(name: str, **arrays):
    np.savez_compressed(OUT / name, **arrays)
    print(f"wrote {name}.npz  ({(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB)")


def reynolds_case(name, mesh, fields, tmp, checkpoint=False):
    stem = "chk" if checkpoint else "plt_cnt"
    p = tmp / f"{name}_hdf5_{stem}_0000"
    synth.write_flash_file(p, mesh, fields, checkpoint=checkpoint)
    radius, stress, means = rh.ref_reynolds_stress(p, 0)
    out = {f"in_{k}": v for k, v in fields.items()}
    out.update({f"mesh_{k}": v for k, v in mesh_arrays(mesh).items()})
    out["radius"] = radius
    out.update({f"stress_{k}": v for k, v in stress.items()})
    out.update({f"mean_{k}": v for k, v in means.items()})
    # axis y / z: the reference's raxis != 0 returns the x-profile; the documented meaning is
    # pinned by running raxis=0 on the axis-permuted file (single-block meshes only)
    if mesh.nblocks == 1:
        for axis, perm in ((1, (0, 2, 1)), (2, (2, 1, 0))):  # [z][y][x] -> bring `axis` to the x slot
            pf = {k: np.ascontiguousarray(np.transpose(v, perm)) for k, v in fields.items()}
            b = mesh.bounds.copy()
            b[[0, axis]] = b[[axis, 0]]
            pm = synth.single_block_mesh(pf["dens"].shape, b)
            pp = tmp / f"{name}_perm{axis}_hdf5_{stem}_0000"
            synth.write_flash_file(pp, pm, pf, checkpoint=checkpoint)
            r, s, m = rh.ref_reynolds_stress(pp, 0)
            out[f"axis{axis}_radius"] = r
            for k, v in s.items():
                out[f"axis{axis}_stress_{k}"] = v
            for k, v in m.items():
                out[f"axis{axis}_mean_{k}"] = v
    save(name, **out)


def uniform_analysis_cases(tmp):
    """G6: fractal_dimension and structure_functions of the reference on uniform 3-D files.
    structure_functions draws from np.random's global state: the seed used is stored with the vectors."""
    _, RefUniform, _ = rh.ref_modules()
    for n, dtype in ((16, np.float32), (32, np.float64)):
        shape = (n, n, n)
        f = synth.uniform_fields(shape, names=FIELDS, dtype=dtype, seed=4000 + n)
        pu = tmp / f"g6_{n}_hdf5_uniform_0000"
        synth.write_flash_file(pu, synth.single_block_mesh(shape, ((0.0, 2.0), (-1.0, 1.0), (0.0, 1.0))), f, uniform3d=True,
                               checkpoint=dtype is np.float64)
        m = RefUniform(str(pu))
        m.load()
        o = {f"in_{k}": v for k, v in f.items()}
        o["bounds"] = np.array([[0.0, 2.0], [-1.0, 1.0], [0.0, 1.0]])
        cases = (("velx", 0.05), ("dens", 1.25), ("vely", float(f["vely"][n // 2, n // 3, n // 4])))
        o["fd_fields"] = np.array([c[0] for c in cases])
        o["fd_contours"] = np.array([c[1] for c in cases])
        for i, (field, contour) in enumerate(cases):
            res = m.fractal_dimension(field, contour)[field][f"{contour}"]
            o[f"fd{i}"] = np.array([res[k] for k in ("average fractal dimension", "slope", "R2", "curve")])
        for tag, kw in (("log", dict(num_seps=5, num_points=300, sep_bounds=[0.02, 0.6], log_scale=True, anistropic=False)),
                        ("lin_aniso", dict(num_seps=4, num_points=257, sep_bounds=[0.1, 1.3], log_scale=False, anistropic=True))):
            np.random.seed(977 + n)
            sf = m.structure_functions(**kw)
            o[f"sf_{tag}_seed"] = np.array(977 + n)
            o[f"sf_{tag}_separations"] = np.array(sf["separations"])
            o[f"sf_{tag}_longitudinal"] = np.stack([sf["longitudinal"][f"{k}"] for k in range(1, 11)])
            o[f"sf_{tag}_transverse"] = np.stack([sf["transverse"][f"{k}"] for k in range(1, 11)])
        save(f"g6_uniform_analysis_{n}", **o)


def slice_cases(tmp):
    """G7: slice_integral / slice_average of the reference (axis 0) on an AMR plt file."""
    RefAMR, _, _ = rh.ref_modules()
    mesh = synth.octree_mesh((2, 1, 2), (4, 8, 4), 3, seed=71, p_refine=0.5, bounds=((0.0, 2.0), (-1.0, 1.0), (0.0, 1.0)))
    fields = synth.block_fields(mesh, names=FIELDS, dtype=np.float32, seed=71)
    p = tmp / "g7_hdf5_plt_cnt_0000"
    synth.write_flash_file(p, mesh, fields)
    m = RefAMR(str(p))
    m.load()
    out = {f"in_{k}": v for k, v in fields.items()}
    out.update({f"mesh_{k}": v for k, v in mesh_arrays(mesh).items()})
    out["span"], out["integral_dens"] = (np.array(v) for v in m.slice_integral("dens", axis=0))
    out["average_velx"] = np.array(m.slice_average("velx", axis=0)[1])
    save("g7_slice_integral", **out)


def main():
    if not rh.reference_available():
        raise SystemExit("the reference is not mounted; goldens can only be generated in the build container")
    tmp = Path(tempfile.mkdtemp(prefix="fava_golden_"))
    if "g6" in sys.argv[1:] or "g7" in sys.argv[1:]:  # only the vectors added after the first batch
        if "g6" in sys.argv[1:]:
            uniform_analysis_cases(tmp)
        if "g7" in sys.argv[1:]:
            slice_cases(tmp)
        return

    # G1: config 1 in miniature — uniform single-block plt (f32), 32x24x16 cells, non-cubic extent
    shape = (16, 24, 32)
    mesh = synth.single_block_mesh(shape, ((0.0, 2.0), (0.0, 1.0), (-1.0, 1.0)))
    fields = synth.uniform_fields(shape, names=FIELDS, dtype=np.float32, seed=1234)
    reynolds_case("g1_uniform_plt_f32", mesh, fields, tmp)

    # G1c: same through a checkpoint (f64) file with a large mean flow (pivot stress test)
    fields64 = synth.uniform_fields(shape, names=FIELDS, dtype=np.float64, seed=99, u0=10.0)
    reynolds_case("g1_uniform_chk_f64", mesh, fields64, tmp, checkpoint=True)

    # G2: multi-block single level (C5-like layout in miniature): 4x2x2 blocks of 8^3
    mb = synth.multiblock_mesh((4, 2, 2), (8, 8, 8))
    full = synth.uniform_fields((16, 16, 32), names=FIELDS, dtype=np.float32, seed=5)
    reynolds_case("g2_multiblock_plt_f32", mb, {k: synth.blocks_from_uniform(mb, v) for k, v in full.items()}, tmp)

    # G3: octree AMR, 4^3 blocks over 3 levels -> 32^3 finest (config 2 in miniature)
    amr = synth.octree_mesh((2, 2, 2), (4, 4, 4), 3, seed=3, p_refine=0.3)
    afields = synth.block_fields(amr, names=FIELDS, dtype=np.float32, seed=1234)
    reynolds_case("g3_amr_plt_f32", amr, afields, tmp)

    # G4: from_amr on the G3 file: whole domain, a true sub-box, refine_level=2 (both), out-of-domain
    p = tmp / "g4_hdf5_plt_cnt_0000"
    synth.write_flash_file(p, amr, afields)
    out = {f"in_{k}": v for k, v in afields.items()}
    out.update({f"mesh_{k}": v for k, v in mesh_arrays(amr).items()})
    cases = {
        "whole": (np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), -1),
        "box": (np.array([[0.25, 0.75], [0.125, 0.5], [0.3, 0.9]]), -1),
        "box_l2": (np.array([[0.25, 0.75], [0.125, 0.5], [0.3, 0.9]]), 2),
        "whole_l2": (np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), 2),
        "whole_l9": (np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), 9),
    }
    from fava_b200 import h5lite

    for tag, (sd, lvl) in cases.items():
        uni = tmp / f"g4_{tag}_hdf5_uniform_0000"
        m, res = rh.ref_from_amr(p, sd, lvl, fields=("dens", "velz"), filename=uni)
        out[f"{tag}_sd"] = sd
        out[f"{tag}_level"] = np.array(lvl)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v  # float64 [NX][NY][NZ] (reference in-memory layout)
        out[f"{tag}_bounds"] = np.array([[m.xmin, m.xmax], [m.ymin, m.ymax], [m.zmin, m.zmax]], dtype=np.float64)
        with h5lite.File(uni) as fh:  # the f32 payload and the metadata quirks of the written file (A10)
            out[f"{tag}_file_dens"] = fh["dens"][()]
            out[f"{tag}_file_bbox"] = fh["bounding box"][()]
            out[f"{tag}_file_blocksize_shape"] = np.array(fh["block size"].shape)
            out[f"{tag}_file_gid_shape"] = np.array(fh["gid"].shape)
            out[f"{tag}_file_whichchild_shape"] = np.array(fh["which child"].shape)
            isc = fh["integer scalars"]
            names = np.char.strip(isc[:, "name"].astype(str))
            vals = isc[:, "value"]
            out[f"{tag}_file_nxb_nyb_nzb"] = np.array([vals[list(names).index(k)] for k in ("nxb", "nyb", "nzb")])
    m, res = rh.ref_from_amr(p, np.array([[0.25, 1.5], [0.1, 0.5], [0.1, 0.5]]), -1, fields=("dens",), filename=tmp / "x")
    out["outside_is_none"] = np.array(res is None)
    save("g4_from_amr", **out)

    # G5: kinetic_energy_spectra on FAVA-style uniform files (3-D datasets), 16^3 f32 and 32^3 f32
    for n in (16, 32):
        shape = (n, n, n)
        f = synth.uniform_fields(shape, names=FIELDS, dtype=np.float32, seed=1234 + n)
        pu = tmp / f"g5_{n}_hdf5_uniform_0000"
        synth.write_flash_file(pu, synth.single_block_mesh(shape), f, uniform3d=True)
        sp = rh.ref_kinetic_energy_spectra(pu)
        o = {f"in_{k}": v for k, v in f.items()}
        o.update({f"spec_{k}": v for k, v in sp.items()})
        save(f"g5_spectrum_{n}", **o)

    uniform_analysis_cases(tmp)
    slice_cases(tmp)


if __name__ == "__main__":
    main()
