"""CPU: the NumPy oracle (oracle/fava_oracle.py) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  This is the pin that lets the GPU tests trust the oracle."""
import numpy as np
import pytest

from oracle import fava_oracle as orc
from tests._util import (FIELDS, STRESS, golden_fields, golden_mesh, load_golden, maxnorm_close, oracle_data,
                         oracle_geom)


@pytest.mark.parametrize("name,bbox_dtype", [
    ("g1_uniform_plt_f32", np.float32),
    ("g1_uniform_chk_f64", np.float64),
    ("g2_multiblock_plt_f32", np.float32),
    ("g3_amr_plt_f32", np.float32),
])
def test_reynolds_stress_axis0_bit_exact(name, bbox_dtype):
    g = load_golden(name)
    mesh = golden_mesh(g)
    radius, stress, means = orc.reynolds_stress(oracle_geom(mesh, bbox_dtype), oracle_data(golden_fields(g)), axis=0)
    assert np.array_equal(radius, g["radius"])
    for k in STRESS:
        assert np.array_equal(stress[k], g[f"stress_{k}"]), k
    for k in FIELDS:
        assert np.array_equal(means[k], g[f"mean_{k}"]), k


@pytest.mark.parametrize("name,bbox_dtype", [("g1_uniform_plt_f32", np.float32), ("g1_uniform_chk_f64", np.float64)])
@pytest.mark.parametrize("axis", [1, 2])
def test_reynolds_stress_axis_yz_equals_reference_on_permuted_file(name, bbox_dtype, axis):
    """The oracle's axis=1/2 == the reference's raxis=0 on the axis-permuted file (same arithmetic, the
    summation order inside a plane differs => 1e-13, not bit-exact)."""
    g = load_golden(name)
    mesh = golden_mesh(g)
    radius, stress, means = orc.reynolds_stress(oracle_geom(mesh, bbox_dtype), oracle_data(golden_fields(g)), axis=axis)
    maxnorm_close(radius, g[f"axis{axis}_radius"], 1e-15, "radius")
    for k in STRESS:
        maxnorm_close(stress[k], g[f"axis{axis}_stress_{k}"], 1e-13, k)
    for k in FIELDS:
        maxnorm_close(means[k], g[f"axis{axis}_mean_{k}"], 1e-13, k)


@pytest.mark.parametrize("tag", ["whole", "box", "box_l2", "whole_l2", "whole_l9"])
def test_from_amr_bit_exact(tag):
    g = load_golden("g4_from_amr")
    mesh = golden_mesh(g)
    geom = oracle_geom(mesh)
    data = oracle_data(golden_fields(g, ("dens", "velz")))
    plan = orc.from_amr_plan(geom, g[f"{tag}_sd"], int(g[f"{tag}_level"]))
    assert not plan.outside
    for k in ("dens", "velz"):
        got = orc.from_amr_gather(geom, plan, data[k])
        assert got.dtype == np.float64 and np.array_equal(got, g[f"{tag}_{k}"]), (tag, k)
    # the uniform file the reference wrote: f32 payload in [z][y][x] order, quirky metadata shapes
    assert np.array_equal(g[f"{tag}_file_dens"], np.swapaxes(g[f"{tag}_dens"], 0, 2).astype(np.float32))
    assert tuple(g[f"{tag}_file_nxb_nyb_nzb"]) == g[f"{tag}_dens"].shape
    assert tuple(g[f"{tag}_file_blocksize_shape"]) == (1, 3, 3)
    assert tuple(g[f"{tag}_file_gid_shape"]) == (15,)
    assert tuple(g[f"{tag}_file_whichchild_shape"]) == (mesh.nblocks,)
    maxnorm_close(plan.refdom_bound_box, g[f"{tag}_bounds"], 1e-15, "refined-domain bounds")


def test_from_amr_dict_restatement_matches_vectorised():
    """The literal per-cell dict restatement (_flash.py:1262-1314) == the np.repeat form, small case."""
    g = load_golden("g4_from_amr")
    mesh = golden_mesh(g)
    geom = oracle_geom(mesh)
    data = oracle_data(golden_fields(g, ("dens",)))
    plan = orc.from_amr_plan(geom, g["box_l2_sd"], 2)
    assert np.array_equal(orc.from_amr_gather_dict(geom, plan, data["dens"]), g["box_l2_dens"])


def test_from_amr_outside_domain_is_silent_none():
    g = load_golden("g4_from_amr")
    assert bool(g["outside_is_none"])
    plan = orc.from_amr_plan(oracle_geom(golden_mesh(g)), np.array([[0.25, 1.5], [0.1, 0.5], [0.1, 0.5]]), -1)
    assert plan.outside


def test_from_amr_zero_in_every_row_means_whole_domain():
    """subdomain_flag = any(0 not in row) (_flash.py:965): a box with a literal 0 in every row is ignored."""
    g = load_golden("g4_from_amr")
    plan = orc.from_amr_plan(oracle_geom(golden_mesh(g)), np.array([[0.0, 0.5], [0.0, 0.5], [0.0, 0.5]]), -1)
    assert not plan.subdomain_flag and tuple(plan.total_cells) == (32, 32, 32)


@pytest.mark.parametrize("n", [16, 32])
@pytest.mark.parametrize("use_scipy", [True, False])
def test_kinetic_energy_spectra_bit_exact(n, use_scipy):
    g = load_golden(f"g5_spectrum_{n}")
    data = {k: orc.load_like_reference(v) for k, v in golden_fields(g).items()}
    sp = orc.kinetic_energy_spectra(data, (n, n, n), use_scipy=use_scipy)
    assert list(sp) == ["k", "total", "longitudinal", "transverse"]
    for k, v in sp.items():
        assert np.array_equal(v, g[f"spec_{k}"], equal_nan=True), k


def test_kinetic_energy_single_mode_known_answer():
    """u_x = cos(2 pi m x), rho = 1 => `total` lives in shell m only:
    the two points (+-m,0,0) carry 1/2 |1/2|^2 each; mean over the shell's count x 4 pi m^2."""
    n, m = 16, 3
    x = np.arange(n) / n
    ux = np.broadcast_to(np.cos(2 * np.pi * m * x)[:, None, None], (n, n, n)).copy()
    data = {"dens": np.ones((n, n, n)), "velx": ux, "vely": np.zeros((n, n, n)), "velz": np.zeros((n, n, n))}
    sp = orc.kinetic_energy_spectra(data, (n, n, n), use_scipy=False)
    k = np.arange(-n // 2, n // 2)
    kk = np.sqrt(k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2)
    count = np.sum((kk >= m - 0.5) & (kk < m + 0.5))
    expect = np.zeros(n // 2 - 1)
    expect[m] = 4 * np.pi * m**2 * (2 * 0.5 * 0.25) / count
    maxnorm_close(sp["total"], expect, 1e-13, "single-mode total")


def test_slice_integral_and_average_bit_exact():
    """§8f rank 1: oracle == the reference's slice_integral / slice_average (axis 0) on the G7 AMR file."""
    g = load_golden("g7_slice_integral")
    mesh = golden_mesh(g)
    geom, data = oracle_geom(mesh), oracle_data(golden_fields(g))
    span, alp = orc.slice_integral(geom, data["dens"], 0)
    assert np.array_equal(span, g["span"]) and np.array_equal(alp, g["integral_dens"])
    assert np.array_equal(orc.slice_average(geom, data["velx"], 0)[1], g["average_velx"])


def test_transposed_projection_equals_the_pointwise_form_shell_by_shell():
    """The identity the binning kernel rests on (csrc/spectrum.cu): the reference sums |sum_n k_n u^_n(rev k)|^2 / |k|^2 over a
    shell (`ffts[n].T`, FlashUniform.py:281); rev is a bijection of the cube that keeps |k|, so the shell sums equal those of
    |k_z u^_x(k) + k_y u^_y(k) + k_x u^_z(k)|^2 / |k|^2 - every term from the values stored at k alone."""
    rng = np.random.default_rng(11)
    n = 16
    dens = 1.0 + 0.5 * rng.random((n, n, n))
    vel = [rng.standard_normal((n, n, n)) for _ in range(3)]
    kk = np.linspace(-n // 2, n // 2 - 1, n)
    k = np.array(np.meshgrid(kk, kk, kk, indexing="ij"))
    kabs = np.sqrt((k**2).sum(axis=0))
    ffts = np.array([np.fft.fftshift(np.fft.fftn(np.sqrt(dens) * v, norm="forward")) for v in vel])
    ref = np.zeros((n, n, n), dtype=np.complex128)
    for i in range(3):
        ref += k[i] * ffts[i].T  # the reference's line
    ref = np.abs(ref / np.maximum(kabs, 1e-99)) ** 2
    mine = np.abs((k[2] * ffts[0] + k[1] * ffts[1] + k[0] * ffts[2]) / np.maximum(kabs, 1e-99)) ** 2
    assert np.allclose(np.sort(ref.ravel()), np.sort(mine.ravel()), rtol=1e-12, atol=1e-300)  # the same multiset of terms
    shell = np.floor(kabs + 0.5).astype(int).ravel()
    a = np.bincount(shell, weights=ref.ravel())
    b = np.bincount(shell, weights=mine.ravel())
    assert np.max(np.abs(a - b)) <= 1e-14 * np.max(np.abs(a))
