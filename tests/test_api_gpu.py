"""GPU parity through the reference-shaped Python API (fava_b200.mesh.FLASH / FlashUniform / fava.flash),
i.e. file -> h5lite -> pinned staging -> libfava_b200 kernels (C ABI) -> NumPy results, against
 (i) the golden vectors produced by the unmodified reference, (ii) the NumPy oracle on larger cases.

Tolerances (BASELINE.json north_star): bit-exact for prolongation / indexing; fp64 profiles and spectra
within 1e-12 relative in max-norm per output array.
"""
import numpy as np
import pytest

from fava_b200 import h5lite, synth
from oracle import fava_oracle as orc
from tests._util import (FIELDS, RTOL, STRESS, golden_fields, golden_mesh, load_golden, maxnorm_close, oracle_data,
                         oracle_geom)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fava(cuda_device):
    import fava_b200

    return fava_b200


def write_golden_file(tmp_path, g, stem="plt_cnt", checkpoint=False, **kw):
    mesh = golden_mesh(g)
    path = tmp_path / f"case_hdf5_{stem}_0000"
    synth.write_flash_file(path, mesh, golden_fields(g), checkpoint=checkpoint, **kw)
    return path, mesh


@pytest.mark.parametrize("name,chk", [("g1_uniform_plt_f32", False), ("g1_uniform_chk_f64", True),
                                      ("g2_multiblock_plt_f32", False), ("g3_amr_plt_f32", False)])
def test_reynolds_stress_axis0_vs_reference_golden(fava, tmp_path, name, chk):
    g = load_golden(name)
    path, _ = write_golden_file(tmp_path, g, "chk" if chk else "plt_cnt", checkpoint=chk)
    m = fava.mesh.FLASH(path)
    m.load()
    radius, stress, means = m.reynolds_stress(raxis=0)
    assert np.array_equal(radius, g["radius"])
    assert list(stress) == list(STRESS) and list(means) == list(FIELDS)
    for k in STRESS:
        maxnorm_close(stress[k], g[f"stress_{k}"], RTOL, k)
    for k in FIELDS:
        maxnorm_close(means[k], g[f"mean_{k}"], RTOL, k)


@pytest.mark.parametrize("name,chk", [("g1_uniform_plt_f32", False), ("g1_uniform_chk_f64", True)])
@pytest.mark.parametrize("axis", [1, 2])
def test_reynolds_stress_axis_yz_vs_reference_on_permuted_file(fava, tmp_path, name, chk, axis):
    g = load_golden(name)
    path, _ = write_golden_file(tmp_path, g, "chk" if chk else "plt_cnt", checkpoint=chk)
    m = fava.mesh.FLASH(path)
    m.load()
    radius, stress, means = m.reynolds_stress(axis=axis)
    maxnorm_close(radius, g[f"axis{axis}_radius"], 1e-15, "radius")
    for k in STRESS:
        maxnorm_close(stress[k], g[f"axis{axis}_stress_{k}"], RTOL, k)
    for k in FIELDS:
        maxnorm_close(means[k], g[f"axis{axis}_mean_{k}"], RTOL, k)


@pytest.mark.parametrize("name", ["g2_multiblock_plt_f32", "g3_amr_plt_f32"])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_block_profiles_and_favre_vs_oracle(fava, tmp_path, name, axis):
    g = load_golden(name)
    path, mesh = write_golden_file(tmp_path, g)
    geom, data = oracle_geom(mesh), oracle_data(golden_fields(g))
    m = fava.mesh.FLASH(path)
    m.load()
    radius, stress, means = m.reynolds_stress(axis=axis)
    r0, s0, m0 = orc.reynolds_stress(geom, data, axis=axis)
    assert np.array_equal(radius, r0)
    for k in STRESS:
        maxnorm_close(stress[k], s0[k], RTOL, f"{k} axis {axis}")
    for k in FIELDS:
        maxnorm_close(means[k], m0[k], RTOL, f"{k} axis {axis}")
    _, fstress, fmeans = m.favre_stress(axis=axis)
    fm0, fs0 = orc.favre_stress(geom, data, axis=axis)
    for k in STRESS:
        maxnorm_close(fstress[k], fs0[k], RTOL, f"favre {k} axis {axis}")
    for k in ("velx", "vely", "velz"):
        maxnorm_close(fmeans[k], fm0[k], RTOL, f"favre mean {k} axis {axis}")


@pytest.mark.parametrize("shape_blocks", [((6, 4, 10), (2, 3, 1)), ((16, 16, 16), (2, 2, 2)), ((8, 8, 8), (4, 2, 2)),
                                          ((32, 32, 32), (1, 2, 1)), ((4, 4, 4), (2, 2, 2))])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_multiblock_shapes_fast_and_generic_paths(fava, tmp_path, shape_blocks, axis):
    """Power-of-two blocks take the register-resident CTA-per-leaf kernel, others the warp-per-plane one."""
    nb_xyz, nroot = shape_blocks
    mesh = synth.multiblock_mesh(nroot, nb_xyz, ((0.0, 2.0), (0.0, 1.0), (-1.0, 1.0)))
    full_shape = (nroot[2] * nb_xyz[2], nroot[1] * nb_xyz[1], nroot[0] * nb_xyz[0])
    full = synth.uniform_fields(full_shape, names=FIELDS, dtype=np.float32, seed=11, u0=3.0)
    fields = {k: synth.blocks_from_uniform(mesh, v) for k, v in full.items()}
    path = tmp_path / "mb_hdf5_plt_cnt_0003"
    synth.write_flash_file(path, mesh, fields)
    m = fava.mesh.FLASH(path)
    m.load()
    radius, stress, means = m.reynolds_stress(axis=axis)
    # a single-level multi-block mesh has the statistics of the assembled uniform array
    geom = orc.uniform_geom(full_shape[::-1], mesh.bounds, bbox_dtype=np.float32)
    r0, s0, m0 = orc.reynolds_stress(geom, oracle_data(full), axis=axis)
    maxnorm_close(radius, r0, 1e-15, "radius")
    for k in STRESS:
        maxnorm_close(stress[k], s0[k], RTOL, f"{k} axis {axis}")
    for k in FIELDS:
        maxnorm_close(means[k], m0[k], RTOL, f"{k} axis {axis}")


@pytest.mark.parametrize("tag", ["whole", "box", "box_l2", "whole_l2", "whole_l9"])
def test_from_amr_bit_exact_vs_reference_golden(fava, tmp_path, tag):
    g = load_golden("g4_from_amr")
    path, mesh = write_golden_file(tmp_path, g)
    m = fava.mesh.FLASH(path)
    m.load()
    out = tmp_path / f"{tag}_hdf5_uniform_0000"
    res = m.from_amr(subdomain_coords=g[f"{tag}_sd"], refine_level=int(g[f"{tag}_level"]), fields=["dens", "velz"],
                     filename=out)
    assert res is None
    for k in ("dens", "velz"):
        got = m.data(k)
        assert got.dtype == np.float64 and np.array_equal(got, g[f"{tag}_{k}"]), (tag, k)
    assert (m.nblocks, m.nblockx) == (1, 1) and (m.nxb, m.nyb, m.nzb) == g[f"{tag}_dens"].shape
    maxnorm_close(np.array([[m.xmin, m.xmax], [m.ymin, m.ymax], [m.zmin, m.zmax]]), g[f"{tag}_bounds"], 1e-15, "bounds")
    with h5lite.File(out) as fh:  # the written uniform file, including the reference's metadata quirks (A10)
        assert fh["dens"].dtype == np.float32 and np.array_equal(fh["dens"][()], g[f"{tag}_file_dens"])
        assert np.array_equal(fh["bounding box"][()], g[f"{tag}_file_bbox"])
        assert fh["block size"].shape == tuple(g[f"{tag}_file_blocksize_shape"])
        assert fh["gid"].shape == tuple(g[f"{tag}_file_gid_shape"])
        assert fh["which child"].shape == tuple(g[f"{tag}_file_whichchild_shape"])
        isc = fh["integer scalars"]
        vals = dict(zip(np.char.strip(isc[:, "name"].astype(str)), isc[:, "value"]))
        assert tuple(vals[k] for k in ("nxb", "nyb", "nzb")) == tuple(g[f"{tag}_file_nxb_nyb_nzb"])
    # the written file loads as a uniform mesh whose statistics equal the AMR ones
    if tag == "whole":
        u = fava.mesh.FlashUniform(out)
        u.load()
        assert u.data("dens").shape == g["whole_dens"].shape


def test_from_amr_quirks(fava, tmp_path):
    g = load_golden("g4_from_amr")
    path, _ = write_golden_file(tmp_path, g)
    m = fava.mesh.FLASH(path)
    m.load()
    # outside the domain: silent None, mesh untouched (_flash.py:967-977)
    assert m.from_amr(np.array([[0.25, 1.5], [0.1, 0.5], [0.1, 0.5]]), fields=["dens"], filename=tmp_path / "x") is None
    assert m.nblocks == golden_mesh(g).nblocks
    with pytest.raises(TypeError):
        m.from_amr(None, fields=["dens"])
    # a literal 0 in every row => the sub-domain is ignored (:965); lists and long names are accepted
    m.from_amr([[0.0, 0.5], [0.0, 0.5], [0.0, 0.5]], fields=["density"], filename=tmp_path / "y_hdf5_uniform_0000")
    assert np.array_equal(m.data("dens"), g["whole_dens"])


def test_amr_statistics_equal_uniform_statistics_of_prolonged_grid(fava, tmp_path):
    """on a multi-level mesh reynolds_stress == the uniform-grid statistic of the
    from_amr(refine_level=-1) output (K1 block front end vs K3 -> K1 dense)."""
    g = load_golden("g3_amr_plt_f32")
    path, _ = write_golden_file(tmp_path, g)
    a = fava.mesh.FLASH(path)
    a.load()
    ra, sa, ma = a.reynolds_stress(raxis=0)
    b = fava.mesh.FLASH(path)
    b.load()
    b.from_amr(np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), fields=["dens", "velx", "vely", "velz"],
               filename=tmp_path / "u_hdf5_uniform_0000")
    rb, sb, mb = b.reynolds_stress(raxis=0)
    maxnorm_close(rb, ra, 1e-15, "radius")
    for k in STRESS:
        maxnorm_close(sb[k], sa[k], 1e-13, k)
    for k in FIELDS:
        maxnorm_close(mb[k], ma[k], 1e-13, k)


@pytest.mark.parametrize("n", [16, 32])
def test_kinetic_energy_spectra_vs_reference_golden(fava, tmp_path, n):
    g = load_golden(f"g5_spectrum_{n}")
    path = tmp_path / f"sp_hdf5_uniform_{n:04d}"
    synth.write_flash_file(path, synth.single_block_mesh((n, n, n)), golden_fields(g), uniform3d=True)
    u = fava.mesh.FlashUniform(path)
    u.load()
    sp = u.kinetic_energy_spectra()
    assert list(sp) == ["k", "total", "longitudinal", "transverse"]
    assert np.array_equal(sp["k"], g["spec_k"])
    for k in ("total", "longitudinal", "transverse"):
        maxnorm_close(sp[k], g[f"spec_{k}"], RTOL, k)


@pytest.mark.parametrize("n,dtype", [(64, np.float64), (96, np.float32), (128, np.float64), (20, np.float64)])
def test_kinetic_energy_spectra_vs_oracle(fava, tmp_path, n, dtype):
    shape = (n, n, n)
    f = synth.uniform_fields(shape, names=FIELDS, dtype=dtype, seed=321 + n, u0=0.5)
    path = tmp_path / f"sp_hdf5_{'chk' if dtype == np.float64 else 'uniform'}_0001"
    synth.write_flash_file(path, synth.single_block_mesh(shape), f, uniform3d=True, checkpoint=dtype == np.float64)
    u = fava.mesh.FlashUniform(path)
    u.load()
    sp = u.kinetic_energy_spectra()
    ref = orc.kinetic_energy_spectra({k: orc.load_like_reference(v) for k, v in f.items()}, (n, n, n), use_scipy=False)
    for k in ("k", "total", "longitudinal", "transverse"):
        maxnorm_close(sp[k], ref[k], RTOL, f"{k} n={n}")


def test_kinetic_energy_single_mode_and_errors(fava, tmp_path, cuda_device):
    import torch

    from fava_b200 import device

    n, mode = 32, 5
    x = np.arange(n) / n
    ux = np.broadcast_to(np.cos(2 * np.pi * mode * x)[None, None, :], (n, n, n)).copy()
    one, zero = np.ones((n, n, n)), np.zeros((n, n, n))
    t = [torch.from_numpy(a).to(cuda_device) for a in (one, ux, zero, zero)]
    sp = device.ke_spectrum(*t)
    k = np.arange(-n // 2, n // 2)
    kk = np.sqrt(k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2)
    count = np.sum((kk >= mode - 0.5) & (kk < mode + 0.5))
    expect = np.zeros(n // 2 - 1)
    expect[mode] = 4 * np.pi * mode**2 * (2 * 0.5 * 0.25) / count
    maxnorm_close(sp["total"], expect, 1e-13, "single-mode total")
    with pytest.raises(ValueError):
        bad = [torch.zeros((8, 8, 16), dtype=torch.float64, device=cuda_device)] * 4
        device.ke_spectrum(*bad)
    with pytest.raises(RuntimeError):
        odd = [torch.zeros((9, 9, 9), dtype=torch.float64, device=cuda_device)] * 4
        device.ke_spectrum(*odd)


@pytest.mark.parametrize("name", ["g1_uniform_plt_f32", "g3_amr_plt_f32"])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_slice_integral_and_average_vs_oracle(fava, tmp_path, name, axis):
    g = load_golden(name)
    path, mesh = write_golden_file(tmp_path, g)
    geom, data = oracle_geom(mesh), oracle_data(golden_fields(g))
    m = fava.mesh.FLASH(path)
    m.load()
    span, alp = m.slice_integral("dens", axis=axis)
    s0, a0 = orc.slice_integral(geom, data["dens"], axis)
    assert np.array_equal(span, s0)
    maxnorm_close(alp, a0, RTOL, "slice_integral")
    _, avg = m.slice_average("velocity-x", axis=axis)
    _, v0 = orc.slice_average(geom, data["velx"], axis)
    maxnorm_close(avg, v0, RTOL, "slice_average")


def test_model_entry_point_and_registries(fava, tmp_path):
    """`fava.flash(dir).load(file_type=...)` + analysis methods registered on Model (README.rst:9-62)."""
    g = load_golden("g1_uniform_plt_f32")
    write_golden_file(tmp_path, g)
    gs = load_golden("g5_spectrum_16")
    synth.write_flash_file(tmp_path / "case_hdf5_uniform_0007", synth.single_block_mesh((16, 16, 16)), golden_fields(gs),
                           uniform3d=True)
    model = fava.flash(tmp_path)
    assert model.nfiles(file_type="plt") == 1 and model.nfiles(file_type="uni") == 1
    assert list(model.uni_files["by number"]) == [7]
    model.load(file_index=0, file_type="plt")
    assert type(model.mesh).__name__ == "FLASH" and model.mesh.mesh_type == "FLASH"
    radius, stress, means = model.reynolds_stress(axis=0)
    maxnorm_close(stress["Rxy"], g["stress_Rxy"], RTOL, "Rxy via model")
    model.load(file_number=7, file_type="uni")
    sp = model.kinetic_energy_spectra()
    maxnorm_close(sp["total"], gs["spec_total"], RTOL, "spectrum via model")
    assert {"FLASH", "FlashUniform", "Structured", "Unstructured"} <= set(fava.Model.mesh_names())
    with pytest.raises(ValueError):
        model.load(file_index=0, file_type="plt")
        model.reynolds_stress(axis=3)


def test_data_returns_reference_layout(fava, tmp_path):
    g = load_golden("g2_multiblock_plt_f32")
    path, _ = write_golden_file(tmp_path, g)
    m = fava.mesh.FLASH(path)
    m.load()
    m.load_data(["dens"])
    ref = orc.load_like_reference(g["in_dens"])
    got = m.data("density")
    assert got.dtype == np.float64 and np.array_equal(got, ref)
    assert m.data("no such field") is None


@pytest.mark.parametrize("n,dtype", [(256, np.float64), (256, np.float32), (512, np.float32), (512, np.float64), (1024, np.float64),
                                     (2048, np.float32)])
def test_hand_written_transform_matches_torch_fft(cuda_device, n, dtype):
    """Power-of-two grids take the hand-written passes (csrc/fft.cu): the fused weighting + x pass, the y pass and the
    pruned z pass, each compared with torch.fft (cuFFT, test reference only) on the elements a bin can read."""
    import torch

    from fava_b200 import device

    assert device.fft_native_supported(n) and not device.fft_native_supported(96) and not device.fft_native_supported(128)
    assert device.spectral_pitch(n) == n // 2 and device.spectral_pitch(96) == 49
    nz = 8
    g = torch.Generator(device=cuda_device)
    g.manual_seed(n)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    rho = (1.0 + 0.5 * torch.rand((nz, n, n), generator=g, device=cuda_device, dtype=torch.float64)).to(tdt)
    u = [(torch.randn((nz, n, n), generator=g, device=cuda_device, dtype=torch.float64) + 0.3 * i).to(tdt) for i in range(3)]
    pitch = n // 2
    out = [torch.zeros((nz, n, pitch), dtype=torch.complex128, device=cuda_device) for _ in range(3)]
    device.ke_transform_x(rho, *u, *[o.data_ptr() for o in out])
    kmax2 = n * n // 4 - 3 * n // 2 + 2
    k = torch.arange(n, device=cuda_device)
    wn = torch.where(k < n // 2, k, k - n)
    kx0 = (torch.arange(pitch, device=cuda_device) // (8192 // n)) * (8192 // n)  # first column of the column tile
    for c in range(3):
        w = torch.sqrt(rho.double()) * u[c].double()
        ref = torch.fft.rfft(w, dim=-1)[..., :pitch]
        maxnorm_close(torch.view_as_real(out[c]).cpu().numpy(), torch.view_as_real(ref).cpu().numpy(), 1e-13, f"x pass {c}")
        device.ke_transform_y(out[c].data_ptr(), nz, n, cuda_device)
        ref2 = torch.fft.fft(ref, dim=1)
        keep = (wn[:, None] ** 2 + kx0[None, :] ** 2) <= kmax2  # rows the pruned y pass writes
        got = torch.where(keep[None], out[c], torch.zeros_like(out[c]))
        want = torch.where(keep[None], ref2, torch.zeros_like(ref2))
        maxnorm_close(torch.view_as_real(got).cpu().numpy(), torch.view_as_real(want).cpu().numpy(), 1e-13, f"y pass {c}")
    # z pass on a ky-pencil [n (z)][nyl][pitch] with a ky map: rows 3, n-3, 40 and a padding row
    ky_rows = torch.tensor([3, n - 3, 40, -1], dtype=torch.int32, device=cuda_device)
    z = (torch.randn((n, 4, pitch), generator=g, device=cuda_device, dtype=torch.float64)
         + 1j * torch.randn((n, 4, pitch), generator=g, device=cuda_device, dtype=torch.float64))
    ref3 = torch.fft.fft(z, dim=0)
    device.ke_transform_z(z.data_ptr(), n, 4, ky_rows, cuda_device)
    ky = torch.tensor([3, -3, 40, 10**6], device=cuda_device)
    keep = (wn[:, None, None] ** 2 + ky[None, :, None] ** 2 + kx0[None, None, :] ** 2) <= kmax2
    maxnorm_close(torch.view_as_real(torch.where(keep, z, torch.zeros_like(z))).cpu().numpy(),
                  torch.view_as_real(torch.where(keep, ref3, torch.zeros_like(ref3))).cpu().numpy(), 1e-13, "z pass")


@pytest.mark.parametrize("n", [256, 96])
def test_spectrum_both_transform_paths_against_the_oracle(cuda_device, n):
    """the oracle at 256^3 (hand-written transform path) and at 96^3 (cuFFT path, not a power of two)."""
    import torch

    from fava_b200 import device

    full = synth.uniform_fields((n, n, n), names=FIELDS, seed=31 + n, u0=1.5)
    want = orc.kinetic_energy_spectra({k: orc.load_like_reference(v) for k, v in full.items()}, (n, n, n))
    t = [torch.from_numpy(full[k].copy()).to(cuda_device) for k in FIELDS]
    got = device.ke_spectrum(*t)
    for key in ("k", "total", "longitudinal", "transverse"):
        maxnorm_close(got[key], want[key], RTOL, f"{key} n={n}")
    again = device.ke_spectrum(*t)
    assert all(np.array_equal(got[q], again[q]) for q in got)
    t32 = [v.float() for v in t]  # plt files hold f32
    full32 = {k: v.astype(np.float32) for k, v in full.items()}
    want32 = orc.kinetic_energy_spectra({k: orc.load_like_reference(v) for k, v in full32.items()}, (n, n, n))
    got32 = device.ke_spectrum(*t32)
    for key in ("total", "longitudinal", "transverse"):
        maxnorm_close(got32[key], want32[key], RTOL, f"{key} n={n} f32")


def test_staging_file_and_host_paths_are_byte_exact(cuda_device, tmp_path):
    """fava_stage_h2d (pread -> pinned ring -> async H2D) and fava_stage_host_h2d deliver the dataset's bytes
    unchanged, for sizes around the 16 MiB chunk / 2 MiB slice boundaries and odd tails."""
    import torch

    from fava_b200 import device

    rng = np.random.default_rng(3)
    for nfloat in (1, 1000, (2 << 20) // 4 + 3, (16 << 20) // 4, (40 << 20) // 4 + 17):
        a = rng.random(nfloat, dtype=np.float32)
        path = tmp_path / f"s{nfloat}_hdf5_plt_cnt_0000"
        with h5lite.File(path, "w") as f:
            f.create_dataset("dens", data=a)
        with h5lite.File(path) as f:
            off, nbytes = f["dens"].extent()
        out = torch.empty(nfloat, dtype=torch.float32, device=cuda_device)
        device.stage_file(path, off, nbytes, out)
        assert np.array_equal(out.cpu().numpy(), a)
        out2 = torch.zeros(nfloat, dtype=torch.float32, device=cuda_device)
        device.stage_host(a, out2)
        assert torch.equal(out, out2)
    with pytest.raises(RuntimeError):
        device.stage_file(tmp_path / "missing", 0, 16, torch.empty(4, dtype=torch.float32, device=cuda_device))
    with pytest.raises(RuntimeError):  # dataset shorter than requested
        device.stage_file(path, off, nbytes + (1 << 20), torch.empty(nfloat + (1 << 18), dtype=torch.float32, device=cuda_device))


def test_reynolds_time_series_over_plt_files(fava, tmp_path):
    """BASELINE configs[4] in miniature: the pipeline's per-file loop (reference __main__.py:76-97) over multi-block
    plt files; every file's profile equals the oracle's on the assembled uniform array."""
    from fava_b200 import series

    mesh = synth.multiblock_mesh((2, 2, 2), (16, 16, 16))
    fulls = []
    for i in range(3):
        full = synth.uniform_fields((32, 32, 32), names=FIELDS, dtype=np.float32, seed=100 + i)
        fulls.append(full)
        synth.write_flash_file(tmp_path / f"ts_hdf5_plt_cnt_{i:04d}", mesh,
                               {k: synth.blocks_from_uniform(mesh, v) for k, v in full.items()}, time=0.5 * i)
    model = fava.flash(tmp_path)
    res, timing = series.reynolds_series(model, axis=0)
    assert timing["files"] == 3 and timing["staged_bytes_this_rank"] == 3 * 4 * 32**3 * 4
    for i, (t, radius, stress, means) in enumerate(res):
        assert t == 0.5 * i
        geom = orc.uniform_geom((32, 32, 32), bbox_dtype=np.float32)
        _, s0, m0 = orc.reynolds_stress(geom, oracle_data(fulls[i]), axis=0)
        for k in STRESS:
            maxnorm_close(stress[k], s0[k], RTOL, f"file {i} {k}")


@pytest.mark.parametrize("n", [256, 1024])
def test_y_pass_fused_with_the_exchange_on_one_gpu(cuda_device, n):
    """fava_fft_y_scatter (the y pass whose output rows go straight into the owners' receive buffers) with this GPU
    playing every rank: two 'ranks' own interleaved ky sets in permuted row order, the Nyquist row has no owner; chunked
    calls (z_offset) fill the same buffers as one call.  Checked against torch.fft on the rows the pass must write."""
    import torch

    from fava_b200 import device

    nz, pitch = 8, n // 2
    g = torch.Generator(device=cuda_device)
    g.manual_seed(n + 1)
    src = (torch.randn((nz, n, pitch), generator=g, device=cuda_device, dtype=torch.float64)
           + 1j * torch.randn((nz, n, pitch), generator=g, device=cuda_device, dtype=torch.float64))
    ref = torch.fft.fft(src, dim=1)
    ky = np.arange(n)
    owner = (ky % 2).astype(np.int32)
    owner[n // 2] = -1
    nyl = n // 2
    row = np.zeros(n, dtype=np.int32)
    for r in (0, 1):
        mine = ky[(owner == r)]
        row[mine] = np.arange(mine.size)[::-1]  # permuted: the kernel must go through the table
    me, nz_local = 1, nz  # this GPU acts as rank 1 of 2: its planes land at z slots [nz_local, 2 nz_local)
    recv = [torch.zeros((2 * nz_local, nyl, pitch), dtype=torch.complex128, device=cuda_device) for _ in range(2)]
    peers = torch.tensor([t.data_ptr() for t in recv], dtype=torch.int64, device=cuda_device)
    t_owner, t_row = torch.from_numpy(owner).to(cuda_device), torch.from_numpy(row).to(cuda_device)
    work = src.clone()
    device.fft_y_scatter(work.data_ptr(), n, nz, peers, t_owner, t_row, me, nz_local, nyl, max_ctas=40)
    kmax2 = n * n // 4 - 3 * n // 2 + 2
    wn = np.where(ky < n // 2, ky, ky - n)
    kx0 = (np.arange(pitch) // (8192 // n)) * (8192 // n)
    keep = torch.from_numpy((wn[:, None] ** 2 + kx0[None, :] ** 2) <= kmax2).to(cuda_device)
    for r in (0, 1):
        mine = torch.from_numpy(ky[owner == r]).to(cuda_device)
        rows = torch.from_numpy(row[ky[owner == r]].astype(np.int64)).to(cuda_device)
        got = recv[r][me * nz_local:(me + 1) * nz_local][:, rows, :]
        want = ref[:, mine, :]
        m = keep[mine][None]
        maxnorm_close(torch.view_as_real(torch.where(m, got, torch.zeros_like(got))).cpu().numpy(),
                      torch.view_as_real(torch.where(m, want, torch.zeros_like(want))).cpu().numpy(), 1e-13, f"rank {r}")
        assert float(recv[r][:me * nz_local].abs().max()) == 0.0  # nobody else's z slots were touched
    # the same slab in two chunks (as stats.host_step feeds it) fills the same buffers bit for bit
    again = [torch.zeros_like(t) for t in recv]
    peers2 = torch.tensor([t.data_ptr() for t in again], dtype=torch.int64, device=cuda_device)
    work = src.clone()
    half = nz // 2
    device.fft_y_scatter(work.data_ptr(), n, half, peers2, t_owner, t_row, me, nz_local, nyl, z_offset=0)
    device.fft_y_scatter(work.data_ptr() + half * n * pitch * 16, n, nz - half, peers2, t_owner, t_row, me, nz_local, nyl, z_offset=half)
    for r in (0, 1):
        assert torch.equal(torch.view_as_real(again[r]), torch.view_as_real(recv[r]))
