"""CPU: box-counting fractal dimension and structure functions — the oracle against the golden
vectors of the unmodified reference (tests/golden/g6_*), the oracle's vectorised marking against its literal
per-cell loop on adversarial inputs, the decision rule the kernel uses instead of the reference's division, and the
host helpers of fava_b200/uniform_analysis.py."""
import numpy as np
import pytest

from fava_b200 import uniform_analysis as ua
from oracle import fava_oracle as orc
from tests._util import load_golden

FD_KEYS = ("average fractal dimension", "slope", "R2", "curve")


def adversarial_field(n=16, seed=3, contour=0.5):
    """fp64 field with exact hits of the contour, crossings whose quotient rounds to exactly 1.0 (huge |val|: the
    differences c - val and nb - val coincide after rounding) and flagged cells on the boundary layer."""
    rng = np.random.default_rng(seed)
    d = rng.random((n, n, n))
    d[rng.random(d.shape) < 0.05] = contour  # exact equality, also on the faces
    low = rng.random(d.shape) < 0.03
    d[low] = -1.0e10  # ulp(1e10) ~ 1.9e-6: neighbours in (c, c + 1e-6) give quotient == 1.0 -> the NEIGHBOUR is flagged
    near = rng.random(d.shape) < 0.2
    d[near & ~low] = contour + 1.0e-7 * rng.random(int((near & ~low).sum()))
    return d


@pytest.mark.parametrize("n", [16, 32])
def test_oracle_reproduces_reference_fractal_dimension_bit_exact(n):
    g = load_golden(f"g6_uniform_analysis_{n}")
    for i, (field, contour) in enumerate(zip(g["fd_fields"], g["fd_contours"])):
        d = orc.load_like_reference(g[f"in_{field}"])
        res = orc.fractal_dimension(d, str(field), float(contour))[str(field)][f"{float(contour)}"]
        assert np.array_equal(np.array([res[k] for k in FD_KEYS]), g[f"fd{i}"], equal_nan=True), (field, contour)


@pytest.mark.parametrize("n", [16, 32])
@pytest.mark.parametrize("tag,kw", [
    ("log", dict(num_seps=5, num_points=300, sep_bounds=[0.02, 0.6], log_scale=True, anistropic=False)),
    ("lin_aniso", dict(num_seps=4, num_points=257, sep_bounds=[0.1, 1.3], log_scale=False, anistropic=True)),
])
def test_oracle_reproduces_reference_structure_functions_bit_exact(n, tag, kw):
    g = load_golden(f"g6_uniform_analysis_{n}")
    vel = {k: orc.load_like_reference(g[f"in_{k}"]) for k in ("velx", "vely", "velz")}
    np.random.seed(int(g[f"sf_{tag}_seed"]))
    sf = orc.structure_functions(vel, (n, n, n), g["bounds"], **kw)
    assert np.array_equal(sf["separations"], g[f"sf_{tag}_separations"])
    for o in range(1, 11):
        assert np.array_equal(sf["longitudinal"][f"{o}"], g[f"sf_{tag}_longitudinal"][o - 1]), o
        assert np.array_equal(sf["transverse"][f"{o}"], g[f"sf_{tag}_transverse"][o - 1]), o


def test_vectorised_marking_equals_the_literal_loop_on_adversarial_input():
    for seed in range(4):
        d = adversarial_field(12, seed)
        a, b = orc.fractal_marks(d, 0.5), orc.fractal_marks_loop(d, 0.5)
        assert np.array_equal(a, b)
    # the quotient == 1.0 branch really occurs in this input: some flag sits on a neighbour, not on the low cell
    d = adversarial_field(12, 0)
    inner = d[1:-1, 1:-1, 1:-1]
    nb = d[2:, 1:-1, 1:-1]
    with np.errstate(invalid="ignore"):
        far = (inner < 0.5) & (nb > 0.5) & (np.trunc((0.5 - inner) / (nb - inner)) != 0)
    assert far.any()


def test_quotient_truncates_to_zero_iff_numerator_is_smaller():
    """csrc/fractal.cu decides with (c - val) < (nb - val) instead of int((c - val) / (nb - val)) == 0."""
    rng = np.random.default_rng(0)
    d = np.concatenate([rng.random(200000), 2.0 ** rng.integers(-1000, 1000, 200000).astype(np.float64),
                        np.full(10, 5e-324), rng.random(1000) * 1e-310])
    for steps in (0, 1, 2, 3):
        h = d.copy()
        for _ in range(steps):
            h = np.nextafter(h, 0.0)
        ok = h > 0
        q = h[ok] / d[ok]
        assert np.array_equal(np.trunc(q) == 0, h[ok] < d[ok])
    h = d * rng.random(d.size)
    ok = h > 0
    assert np.array_equal(np.trunc(h[ok] / d[ok]) == 0, h[ok] < d[ok])


def test_box_counts_and_fit_helpers():
    e = np.zeros((8, 8, 8), dtype=np.int8)
    e[0, 0, 0] = e[7, 7, 7] = e[3, 4, 3] = 1
    assert orc.box_counts(e).tolist() == [3, 3, 3, 1]
    assert ua.box_levels((8, 8, 8)) == 4 and ua.box_levels((64, 32, 96)) == 6 and ua.box_levels((48, 48, 48)) == 6
    with pytest.raises(IndexError):
        orc.box_counts(np.zeros((12, 8, 8), dtype=np.int8))
    for counts in ([3, 3, 3, 1], [5483, 3082, 512, 64, 8, 1], [0, 0, 0, 0], [7, 1, 1]):
        a, b = ua.box_count_fit(np.array(counts)), orc.fractal_regression(np.array(counts))
        assert list(a) == list(FD_KEYS)
        for k in FD_KEYS:
            assert np.array_equal(a[k], b[k], equal_nan=True), (counts, k)


def test_tile_plane_ranges_partition_the_grid():
    for nz in (16, 32, 64, 96, 100, 1024):
        for world in (1, 2, 3, 8):
            spans = [ua.tile_plane_range(nz, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nz
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(z0 % ua.TILE == 0 for z0, z1 in spans if z1 > z0)  # ranks beyond the last tile hold nothing
            assert all(z1 % ua.TILE == 0 or z1 == nz for _, z1 in spans)


def test_point_pairs_consume_the_global_stream_like_the_reference():
    bounds = np.array([[0.0, 2.0], [-1.0, 1.0], [0.25, 1.0]])
    np.random.seed(5)
    a = [ua.draw_point_pairs(bounds, s, 1000) for s in (0.01, 0.7, 3.1)]
    state = np.random.random()
    np.random.seed(5)
    b = [orc.structure_function_points(bounds, s, 1000) for s in (0.01, 0.7, 3.1)]
    assert state == np.random.random()
    for (a1, a2), (b1, b2) in zip(a, b):
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
        assert np.all(a2 >= bounds[:, 0]) and np.all(a2 <= bounds[:, 1])


def test_argument_errors_match_the_reference():
    class M:  # the checks below run before any device access
        nCellsVec = np.array([40, 40, 40], dtype=np.int32)
        ndim = 3

    with pytest.raises(ValueError, match="Contours must be either a float or list of floats"):
        ua.fractal_dimension(M(), "dens", 1)  # an int is rejected by the reference too (FlashUniform.py:87-90)
    with pytest.raises(ValueError, match="multiple of 32"):
        ua.fractal_dimension(M(), "dens", 0.5)
    M.ndim = 2
    with pytest.raises(NotImplementedError):
        ua.structure_functions(M())
