"""GPU parity: dense plane-moment kernels (through the C ABI) vs the NumPy oracle.

Tolerance (BASELINE.json north_star): fp64 profiles within 1e-12 relative, max-norm per output array:
max|a-b| <= 1e-12 * max|b|.
"""


import numpy as np
import pytest
import torch

from fava_b200 import synth
from oracle import fava_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def maxnorm_close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    err = np.max(np.abs(a - b))
    assert err <= rtol * scale + 1e-300, f"{what}: max|a-b|={err:.3e} > {rtol:g}*max|b|={scale:.3e}"


def run_gpu(fields, axis, cell_volume, layer_volume, dev):
    import torch

    from fava_b200 import device

    t = {k: torch.from_numpy(v).to(dev) for k, v in fields.items()}
    out = device.plane_profiles(t["dens"], t["velx"], t["vely"], t["velz"], axis, cell_volume, layer_volume)
    return {k: v.cpu().numpy() for k, v in out.items()}


def oracle_profiles(fields, axis, bounds):
    nz, ny, nx = fields["dens"].shape
    geom = orc.uniform_geom((nx, ny, nz), bounds, bbox_dtype=np.float64)
    data = {k: orc.load_like_reference(v)[None, ...] for k, v in fields.items()}
    radius, stress, means = orc.reynolds_stress(geom, data, axis=axis)
    fmeans, favre = orc.favre_stress(geom, data, axis=axis)
    return geom, radius, stress, means, fmeans, favre


def compare(fields, axis, bounds, dev):
    from fava_b200.device import MEAN_KEYS, STRESS_KEYS

    geom, radius, stress, means, fmeans, favre = oracle_profiles(fields, axis, bounds)
    cell_volume = geom.cell_volume_from_level(1)
    db = geom.domain_bounds
    others = [a for a in range(3) if a != axis]
    layer_volume = (db[others[0], 1] - db[others[0], 0]) * (db[others[1], 1] - db[others[1], 0]) * geom.min_delta(axis)
    got = run_gpu(fields, axis, cell_volume, layer_volume, dev)
    for i, k in enumerate(MEAN_KEYS):
        maxnorm_close(got["means"][i], means[k], what=f"mean {k} axis {axis}")
    for i, k in enumerate(STRESS_KEYS):
        maxnorm_close(got["reynolds"][i], stress[k], what=f"reynolds {k} axis {axis}")
        maxnorm_close(got["favre"][i], favre[k], what=f"favre {k} axis {axis}")
    for i, k in enumerate(("velx", "vely", "velz")):
        maxnorm_close(got["favre_means"][i], fmeans[k], what=f"favre mean {k} axis {axis}")


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64, 64), (32, 48, 80), (8, 8, 8)])
def test_dense_profiles_vs_oracle(cuda_device, axis, dtype, shape):
    fields = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"), dtype=dtype, seed=1234)
    compare(fields, axis, ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), cuda_device)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_dense_profiles_large_mean_and_noncubic_extent(cuda_device, axis):
    """mean/rms = 40 exercises the pivot; non-unit extents exercise cell/layer volumes."""
    fields = synth.uniform_fields((48, 40, 56), names=("dens", "velx", "vely", "velz"), u0=10.0, seed=77)
    compare(fields, axis, ((0.0, 2.0), (0.0, 1.0), (-1.0, 1.0)), cuda_device)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_odd_sizes_take_scalar_path(cuda_device, axis):
    fields = synth.uniform_fields((7, 9, 11), names=("dens", "velx", "vely", "velz"), seed=5)
    compare(fields, axis, ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), cuda_device)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_constant_velocity_gives_zero_stress(cuda_device, axis):
    """constant velocity => R_ij = 0 (exactly, thanks to the pivot)."""
    import torch

    from fava_b200 import device

    shape = (32, 32, 32)
    rho = torch.from_numpy(synth.field_slab("dens", shape)).to(cuda_device)
    u = [torch.full(shape, v, dtype=torch.float64, device=cuda_device) for v in (3.0, -2.0, 0.5)]
    out = device.plane_profiles(rho, u[0], u[1], u[2], axis, 1.0 / 32**3, 1.0 / 32)
    assert float(out["reynolds"].abs().max()) == 0.0
    assert float(out["favre"].abs().max()) == 0.0


def test_constant_density_favre_equals_reynolds(cuda_device):
    import torch

    from fava_b200 import device

    shape = (32, 32, 32)
    f = synth.uniform_fields(shape, names=("velx", "vely", "velz"))
    rho = torch.full(shape, 2.0, dtype=torch.float64, device=cuda_device)
    u = [torch.from_numpy(f[k]).to(cuda_device) for k in ("velx", "vely", "velz")]
    out = device.plane_profiles(rho, u[0], u[1], u[2], 0, 1.0 / 32**3, 1.0 / 32)
    r, fv = out["reynolds"].cpu().numpy(), out["favre"].cpu().numpy()
    maxnorm_close(fv, r, rtol=1e-13, what="favre vs reynolds at constant density")


def test_run_to_run_bitwise_determinism(cuda_device):
    import torch

    from fava_b200 import device

    shape = (64, 64, 64)
    f = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"))
    t = [torch.from_numpy(f[k]).to(cuda_device) for k in ("dens", "velx", "vely", "velz")]
    for axis in (0, 1, 2):
        a, _ = device.plane_moments(*t, axis)
        b, _ = device.plane_moments(*t, axis)
        assert torch.equal(a, b)


def test_accumulate_over_z_slabs_matches_one_shot(cuda_device):
    """Chunk-streamed slabs (config 5): moments accumulated slab by slab == whole array (axes x, y)."""
    import torch

    from fava_b200 import device

    shape = (48, 32, 64)
    f = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"))
    t = [torch.from_numpy(f[k]).to(cuda_device) for k in ("dens", "velx", "vely", "velz")]
    for axis in (0, 1):
        whole, piv = device.plane_moments(*t, axis)
        acc = torch.zeros_like(whole)
        for z0 in range(0, 48, 16):
            sl = [x[z0 : z0 + 16].contiguous() for x in t]
            device.plane_moments(*sl, axis, pivots=piv, out=acc, accumulate=True)
        w = whole.cpu().numpy()
        maxnorm_close(acc.cpu().numpy(), w, rtol=1e-13, what=f"slab accumulate axis {axis}")


def test_repivot_is_consistent(cuda_device):
    import torch

    from fava_b200 import device

    shape = (32, 32, 32)
    f = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"), u0=1.0)
    t = [torch.from_numpy(f[k]).to(cuda_device) for k in ("dens", "velx", "vely", "velz")]
    mom, piv = device.plane_moments(*t, 0)
    piv2 = piv + 0.125
    mom2, _ = device.plane_moments(*t, 0, pivots=piv2)
    device.moments_repivot(mom, piv, piv2)
    maxnorm_close(mom.cpu().numpy(), mom2.cpu().numpy(), rtol=1e-13, what="repivot")


def test_bad_arguments_raise(cuda_device):
    import torch

    from fava_b200 import device

    x = torch.zeros((4, 4, 4), dtype=torch.float64, device=cuda_device)
    with pytest.raises(ValueError):
        device.plane_moments(x, x, x, x, 3)
    with pytest.raises(TypeError):
        h = x.to(torch.float16)
        device.plane_moments(h, h, h, h, 0)
    with pytest.raises(ValueError):
        c = x.cpu()
        device.plane_moments(c, c, c, c, 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64, 64), (24, 40, 56), (7, 9, 11)])
def test_fused_xz_pass_matches_single_axis_passes(cuda_device, dtype, shape):
    """fava_plane_moments_xz: x-bins and z-bins from one pass (z from the per-plane column partials, re-pivoted)
    == the dedicated axis-0 and axis-2 passes, after finalisation, to 1e-13; and vs the oracle to 1e-12."""
    import torch

    from fava_b200 import device

    f = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"), dtype=dtype, seed=21, u0=5.0)
    t = [torch.from_numpy(f[k].copy()).to(cuda_device) for k in ("dens", "velx", "vely", "velz")]
    (mx, px), (mz, pz) = device.plane_moments_xz(*t)
    nz, ny, nx = shape
    cv = 1.0 / (nx * ny * nz)
    for axis, mom, piv, n in ((0, mx, px, nx), (2, mz, pz, nz)):
        lv = 1.0 / n
        fused = device.moments_finalize(mom, piv, cv, lv)
        single = device.plane_profiles(*t, axis, cv, lv)
        for k in fused:
            maxnorm_close(fused[k].cpu().numpy(), single[k].cpu().numpy(), rtol=1e-13, what=f"xz {k} axis {axis}")
    compare(f, 0, ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), cuda_device)
    a, _ = device.plane_moments_xz(*t)
    b, _ = device.plane_moments_xz(*t)
    assert torch.equal(a[0], b[0])


@pytest.mark.parametrize("shape,dtype", [((16, 64, 256), torch.float64), ((5, 8, 512), torch.float32), ((40, 136, 768), torch.float64)])
def test_three_axis_pass_equals_per_axis_passes(cuda_device, shape, dtype):
    """fava_plane_moments_xyz (one read of the fields) gives the profiles of the three per-axis calls, for shapes that
    split unevenly over the persistent CTAs, f32 and f64, with a large mean (the pivots matter), bitwise repeatable."""
    from fava_b200 import device

    nz, ny, nx = shape
    g = torch.Generator(device=cuda_device)
    g.manual_seed(nz * 131 + nx)
    f = [(torch.rand(shape, generator=g, device=cuda_device, dtype=torch.float64) + (0.5 if i == 0 else 25.0 * i)).to(dtype)
         for i in range(4)]
    assert device.plane_moments_xyz_supported(shape) and not device.plane_moments_xyz_supported((4, 12, 256))
    assert not device.plane_moments_xyz_supported((4, 8, 384))
    res = device.plane_moments_xyz(*f)
    cv = 1.0 / (nz * ny * nx)
    for axis, (mom, piv) in enumerate(res):
        lv = 1.0 / shape[2 - axis]
        got = device.moments_finalize(mom, piv, cv, lv)
        want = device.plane_profiles(*f, axis, cv, lv)
        for k in want:
            maxnorm_close(got[k].cpu().numpy(), want[k].cpu().numpy(), rtol=1e-13, what=f"xyz {k} axis {axis} {shape}")
    again = device.plane_moments_xyz(*f)
    assert all(torch.equal(a[0], b[0]) for a, b in zip(res, again))


@pytest.mark.parametrize("block,dtype", [((8, 8, 8), torch.float32), ((8, 8, 8), torch.float64), ((16, 8, 8), torch.float32),
                                         ((4, 4, 4), torch.float32)])
def test_small_blocks_streamed_through_the_leaf_ring(cuda_device, block, dtype):
    """Blocks of <= 16 KB take the persistent kernel that streams leaves through shared-memory rings
    (k_block_moments_ring): enough leaves that every group recycles its slots several times, leaves in a shuffled
    order, all three axes, against the dense kernels on the assembled array; bitwise repeatable."""
    from fava_b200 import device

    nzb, nyb, nxb = block
    per = (8, 16, 32)  # blocks along z, y, x
    nz, ny, nx = per[0] * nzb, per[1] * nyb, per[2] * nxb
    nblk = per[0] * per[1] * per[2]
    g = torch.Generator(device=cuda_device)
    g.manual_seed(17 + nzb)
    dense = [(torch.rand((nz, ny, nx), generator=g, device=cuda_device, dtype=torch.float64) + (0.5 if i == 0 else 7.0 * i)).to(dtype)
             for i in range(4)]
    perm = np.random.default_rng(5).permutation(nblk)  # file order of the blocks

    def to_blocks(a):
        b = a.view(per[0], nzb, per[1], nyb, per[2], nxb).permute(0, 2, 4, 1, 3, 5).reshape(nblk, nzb, nyb, nxb)
        out = torch.empty_like(b)
        out[torch.from_numpy(perm).to(cuda_device)] = b  # lattice block i is stored as block perm[i]
        return out.contiguous()

    blocks = [to_blocks(a) for a in dense]
    lattice = np.arange(nblk)
    bz, by, bx = lattice // (per[1] * per[2]), (lattice // per[2]) % per[1], lattice % per[2]
    cv = 1.0 / (nz * ny * nx)
    for axis in (0, 1, 2):
        nbins = (nx, ny, nz)[axis]
        ilo = ((bx * nxb), (by * nyb), (bz * nzb))[axis]
        order = np.random.default_rng(axis).permutation(nblk)  # leaf order of the table
        table = device.leaf_table(perm[order], ilo[order], np.ones(nblk, dtype=np.int64), np.full(nblk, cv))
        mom, piv = device.plane_moments_blocks(*blocks, axis, table, nbins)
        lv = 1.0 / nbins
        got = device.moments_finalize(mom, piv, 1.0, lv)
        want = device.plane_profiles(*dense, axis, cv, lv)
        for k in want:
            maxnorm_close(got[k].cpu().numpy(), want[k].cpu().numpy(), rtol=1e-12, what=f"ring {k} axis {axis} {block}")
        again, _ = device.plane_moments_blocks(*blocks, axis, table, nbins)
        assert torch.equal(mom, again)


def test_leaf_table_uid_keys_the_cached_device_tables(cuda_device):
    """Two immutable tables of the same length but different contents (weights, bins) must not share cached device
    tables: the uid a HostTable carries is what the library's cache is keyed on (fava_plane_moments_blocks_uid)."""
    from fava_b200 import device

    nblk, nb, nbins = 64, 8, 64
    g = torch.Generator(device=cuda_device)
    g.manual_seed(3)
    f = [torch.rand((nblk, nb, nb, nb), generator=g, device=cuda_device, dtype=torch.float32) + 0.5 for _ in range(4)]
    blocks = np.arange(nblk)
    ilo = (blocks % 8) * nb
    ta = device.leaf_table(blocks, ilo, np.ones(nblk, dtype=np.int64), np.full(nblk, 1.0))
    tb = device.leaf_table(blocks, ilo[::-1].copy(), np.ones(nblk, dtype=np.int64), np.full(nblk, 2.0))
    assert ta.uid != tb.uid and not ta.arr.flags.writeable
    with pytest.raises(ValueError):
        ta.arr["vol_frac"][0] = 3.0
    ma, _ = device.plane_moments_blocks(*f, 0, ta, nbins)
    mb, _ = device.plane_moments_blocks(*f, 0, tb, nbins)
    ma2, _ = device.plane_moments_blocks(*f, 0, ta, nbins)
    assert torch.equal(ma, ma2)
    # total density moment: table b weights every leaf twice
    assert abs(float(mb[0].sum()) / float(ma[0].sum()) - 2.0) < 1e-12
    assert not torch.equal(ma[0], 0.5 * mb[0])  # and maps the leaves to other bins
