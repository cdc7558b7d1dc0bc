"""GPU parity of the box-counting fractal dimension and the structure functions through the
reference-shaped API (FlashUniform.fractal_dimension / .structure_functions -> C ABI -> csrc/fractal.cu,
csrc/structure.cu) against the golden vectors of the unmodified reference and the NumPy oracle.

Tolerances: box counts are integers -> bit-exact, and so is the fit computed from them; structure functions are
fp64 sums -> 1e-12 relative in max-norm per array (summation order and pow() differ from NumPy's)."""
import numpy as np
import pytest
import torch

from fava_b200 import synth
from fava_b200 import uniform_analysis as ua
from oracle import fava_oracle as orc
from tests._util import RTOL, load_golden, maxnorm_close
from tests.test_uniform_analysis_cpu import FD_KEYS, adversarial_field

pytestmark = pytest.mark.gpu
NAMES = ("dens", "velx", "vely", "velz")


@pytest.fixture()
def fava(cuda_device):
    import fava_b200

    return fava_b200


def uniform_mesh(fava, tmp_path, fields, bounds=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), f64=False, tag="case"):
    shape = next(iter(fields.values())).shape
    path = tmp_path / f"{tag}_hdf5_uniform_0000"
    synth.write_flash_file(path, synth.single_block_mesh(shape, bounds), fields, uniform3d=True, checkpoint=f64)
    m = fava.mesh.FlashUniform(path)
    m.load()
    return m


def box_counts_gpu(field: np.ndarray, contour: float, splits=None) -> np.ndarray:
    """Counts of every level straight from the two kernels; `splits` = z-ranges handled by separate calls on
    halo-padded sub-buffers, the way several ranks do it."""
    from fava_b200 import device

    nz, ny, nx = field.shape
    dev = torch.device("cuda", 0)
    counts = torch.zeros(32, dtype=torch.int64, device=dev)
    coarse = torch.zeros([(v + 31) // 32 for v in (nz, ny, nx)], dtype=torch.uint8, device=dev)
    for z0, z1 in (splits or [(0, nz)]):
        zf0, zf1 = max(z0 - 1, 0), min(z1 + 1, nz)
        part = torch.from_numpy(np.ascontiguousarray(field[zf0:zf1])).to(dev)
        device.fractal_tiles(part, contour, nz, zf0, z0, z1, counts, coarse)
    nlev = ua.box_levels((nx, ny, nz))
    device.fractal_coarse(coarse, (nz, ny, nx), nlev, counts)
    return counts[:nlev].cpu().numpy()


def oracle_counts(field_zyx: np.ndarray, contour: float) -> np.ndarray:
    return orc.box_counts(orc.fractal_marks(orc.load_like_reference(field_zyx), contour))


@pytest.mark.parametrize("n,f64", [(16, False), (32, True)])
def test_fractal_dimension_vs_reference_golden(fava, tmp_path, n, f64):
    g = load_golden(f"g6_uniform_analysis_{n}")
    m = uniform_mesh(fava, tmp_path, {k: g[f"in_{k}"] for k in NAMES}, g["bounds"], f64)
    for i, (field, contour) in enumerate(zip(g["fd_fields"], g["fd_contours"])):
        out = m.fractal_dimension(str(field), float(contour))
        assert list(out) == [str(field)] and list(out[str(field)]) == [f"{float(contour)}"]
        res = out[str(field)][f"{float(contour)}"]
        assert list(res) == list(FD_KEYS)
        assert np.array_equal(np.array([res[k] for k in FD_KEYS]), g[f"fd{i}"], equal_nan=True), (field, contour)


@pytest.mark.parametrize("shape,dtype", [((64, 64, 64), np.float64), ((96, 32, 64), np.float32), ((128, 128, 128), np.float32),
                                         ((8, 8, 8), np.float64), ((32, 64, 256), np.float64)])
def test_box_counts_bit_exact_vs_oracle(cuda_device, shape, dtype):
    f = synth.uniform_fields(shape, names=("velx", "dens"), dtype=dtype, seed=11 + shape[0])
    for name, contour in (("velx", 0.0), ("velx", 0.3), ("dens", 1.2), ("dens", 9.0)):
        got = box_counts_gpu(f[name], contour)
        assert np.array_equal(got, oracle_counts(f[name], contour)), (shape, name, contour, got)
    assert box_counts_gpu(f["dens"], 9.0).sum() == 0  # nothing reaches the contour: log2(0) = -inf downstream


def test_box_counts_smooth_interface_and_upper_levels(cuda_device):
    """A wrinkled sheet: the counts drop by ~4x per level (dimension ~2), levels >= 6 come from the coarse grid."""
    n = 256
    z, y, x = np.meshgrid(*(np.arange(n, dtype=np.float64),) * 3, indexing="ij")
    field = (x - 0.5 * n - 20.0 * np.sin(2 * np.pi * y / n) * np.cos(4 * np.pi * z / n) - 0.25).astype(np.float32)
    got = box_counts_gpu(field, 0.0)
    assert got.shape == (9,) and np.array_equal(got, oracle_counts(field, 0.0))
    assert got[-1] == 1 and 2.5 < got[0] / got[1] < 5.0


def test_box_counts_adversarial_ties_and_exact_hits(cuda_device):
    for seed in range(3):
        d = adversarial_field(32, seed)  # [i,j,k]; the marking rule is symmetric under the axis swap
        assert np.array_equal(box_counts_gpu(np.ascontiguousarray(d.transpose(2, 1, 0)), 0.5),
                              orc.box_counts(orc.fractal_marks(d, 0.5)))


def test_box_counts_split_over_plane_ranges_equal_one_call(cuda_device):
    f = synth.uniform_fields((128, 64, 192), names=("velx",), dtype=np.float32, seed=5)["velx"]
    whole = box_counts_gpu(f, 0.1)
    assert np.array_equal(whole, box_counts_gpu(f, 0.1, splits=[(0, 32), (32, 96), (96, 128)]))
    assert np.array_equal(whole, oracle_counts(f, 0.1))


def test_fractal_dimension_argument_errors(fava, tmp_path):
    f = synth.uniform_fields((40, 40, 40), names=NAMES, dtype=np.float32, seed=1)
    m = uniform_mesh(fava, tmp_path, f)
    with pytest.raises(ValueError):
        m.fractal_dimension("dens", 0.5)  # 40 is not a multiple of the largest box (32)
    with pytest.raises(ValueError):
        m.fractal_dimension("dens", 1)
    f = synth.uniform_fields((32, 32, 32), names=NAMES, dtype=np.float32, seed=1)
    m = uniform_mesh(fava, tmp_path, f, tag="b")
    with pytest.raises(KeyError):
        m.fractal_dimension("nope", 0.5)
    two = m.fractal_dimension("density", [1.2, 1.3])["density"]  # long names and lists are accepted (superset)
    assert list(two) == ["1.2", "1.3"]
    assert two["1.2"] == m.fractal_dimension("dens", 1.2)["dens"]["1.2"]


@pytest.mark.parametrize("n,f64", [(16, False), (32, True)])
@pytest.mark.parametrize("tag,kw", [
    ("log", dict(num_seps=5, num_points=300, sep_bounds=[0.02, 0.6], log_scale=True, anistropic=False)),
    ("lin_aniso", dict(num_seps=4, num_points=257, sep_bounds=[0.1, 1.3], log_scale=False, anistropic=True)),
])
def test_structure_functions_vs_reference_golden(fava, tmp_path, n, f64, tag, kw):
    g = load_golden(f"g6_uniform_analysis_{n}")
    m = uniform_mesh(fava, tmp_path, {k: g[f"in_{k}"] for k in NAMES}, g["bounds"], f64)
    np.random.seed(int(g[f"sf_{tag}_seed"]))
    sf = m.structure_functions(**kw)
    assert list(sf) == ["transverse", "longitudinal", "separations"]
    assert np.array_equal(sf["separations"], g[f"sf_{tag}_separations"])
    assert list(sf["longitudinal"]) == [f"{o}" for o in range(1, 11)]
    for o in range(1, 11):
        maxnorm_close(sf["longitudinal"][f"{o}"], g[f"sf_{tag}_longitudinal"][o - 1], RTOL, f"long {o}")
        maxnorm_close(sf["transverse"][f"{o}"], g[f"sf_{tag}_transverse"][o - 1], RTOL, f"trans {o}")


def test_structure_functions_vs_oracle_default_sample_size(fava, tmp_path):
    shape = (48, 64, 80)
    bounds = ((0.0, 2.0), (-1.0, 1.0), (0.0, 3.0))
    f = synth.uniform_fields(shape, names=NAMES, dtype=np.float32, seed=21)
    m = uniform_mesh(fava, tmp_path, f, bounds)
    kw = dict(num_seps=7, num_points=10000, sep_bounds=[0.01, 2.5])
    np.random.seed(99)
    got = m.structure_functions(**kw)
    vel = {k: orc.load_like_reference(f[k]) for k in NAMES[1:]}
    np.random.seed(99)
    ref = orc.structure_functions(vel, (80, 64, 48), bounds, **kw)
    for o in range(1, 11):
        maxnorm_close(got["longitudinal"][f"{o}"], ref["longitudinal"][f"{o}"], RTOL, f"long {o}")
        maxnorm_close(got["transverse"][f"{o}"], ref["transverse"][f"{o}"], RTOL, f"trans {o}")
    with pytest.raises(ValueError):  # the reference's default sep_bounds start at 0: np.geomspace refuses
        m.structure_functions()


def test_sf_gather_slabs_sum_to_the_whole_and_out_of_grid_is_flagged(cuda_device):
    from fava_b200 import device

    dev = torch.device("cuda", 0)
    nz, ny, nx = 24, 16, 32
    rng = np.random.default_rng(2)
    vel = [torch.from_numpy(rng.standard_normal((nz, ny, nx)).astype(np.float32)).to(dev) for _ in range(3)]
    lo, cell = np.array([0.0, -1.0, 2.0]), np.array([2.0 / nx, 2.0 / ny, 1.0 / nz])
    pts = rng.random((5000, 3)) * np.array([2.0, 2.0, 1.0]) + lo
    p = torch.from_numpy(pts).to(dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    whole = device.sf_gather(p, *vel, nz, 0, lo, cell, err)
    idx = [np.floor((pts[:, j] - lo[j]) / cell[j]).astype(int) for j in range(3)]
    for c in range(3):
        assert np.array_equal(whole[:, c].cpu().numpy(), vel[c].cpu().numpy().astype(np.float64)[idx[2], idx[1], idx[0]])
    parts = sum(device.sf_gather(p, *[v[a:b].contiguous() for v in vel], nz, a, lo, cell, err) for a, b in ((0, 8), (8, 24)))
    assert torch.equal(parts, whole) and int(err.item()) == 0
    edge = torch.tensor([[2.0, 0.0, 2.5]], dtype=torch.float64, device=dev)  # x == xmax: index nx, IndexError in NumPy
    device.sf_gather(edge, *vel, nz, 0, lo, cell, err)
    assert int(err.item()) == 1
