"""GPU parity at the sizes BASELINE.json names (configs[1..3]), through size-independent properties where the
oracle is too slow or too large:

  C2  8^3-cell blocks over 4 refinement levels -> from_amr(refine_level=-1) to a 256^3 uniform grid: bit-exact
      against the vectorised oracle gather; AMR statistics == statistics of the prolonged grid.
  C3  512^3 uniform fp64 profiles along x/y/z: linearity / known-answer properties + fused-pass consistency.
  C4  (single GPU share) 512^3 spectrum: Parseval-type identity total = longitudinal + transverse and
      sum of shell sums == 0.5 * mean(rho |u|^2) restricted to the binned sphere (checked with a band-limited field).
"""
import numpy as np
import pytest

from fava_b200 import synth
from oracle import fava_oracle as orc
from tests._util import FIELDS, STRESS, maxnorm_close, oracle_data, oracle_geom

pytestmark = pytest.mark.gpu


def test_c2_from_amr_256_cubed_bit_exact_and_statistics_invariant(cuda_device, tmp_path):
    import fava_b200 as fava

    mesh = synth.octree_mesh((4, 4, 4), (8, 8, 8), 4, seed=11, p_refine=0.5)  # 5944 blocks, 5209 leaves
    assert mesh.lmax == 4 and mesh.fine_dims_xyz() == (256, 256, 256)
    fields = synth.block_fields(mesh, names=FIELDS, dtype=np.float32, seed=5)
    path = tmp_path / "c2_hdf5_plt_cnt_0000"
    synth.write_flash_file(path, mesh, fields)
    geom = oracle_geom(mesh)
    whole = np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]])
    plan = orc.from_amr_plan(geom, whole, -1)
    m = fava.mesh.FLASH(path)
    m.load()
    ra, sa, ma = m.reynolds_stress(raxis=0)  # block-list kernels on the AMR mesh
    m.from_amr(whole, refine_level=-1, fields=list(FIELDS), filename=tmp_path / "c2_hdf5_uniform_0000")
    for k in ("dens", "velz"):
        want = orc.from_amr_gather(geom, plan, orc.load_like_reference(fields[k]))
        assert np.array_equal(m.data(k), want), k
    rb, sb, mb = m.reynolds_stress(raxis=0)  # dense kernels on the prolonged 256^3 grid
    for k in STRESS:
        maxnorm_close(sb[k], sa[k], 1e-12, k)
    for k in FIELDS:
        maxnorm_close(mb[k], ma[k], 1e-12, k)
    # a true sub-box with no literal zero, and a coarser target level
    box = np.array([[0.125, 0.6], [0.25, 0.9], [0.3, 0.7]])
    for level in (-1, 3):
        m2 = fava.mesh.FLASH(path)
        m2.load()
        m2.from_amr(box, refine_level=level, fields=["dens"], filename=tmp_path / f"c2b{level}_hdf5_uniform_0000")
        p2 = orc.from_amr_plan(geom, box, level)
        assert np.array_equal(m2.data("dens"), orc.from_amr_gather(geom, p2, orc.load_like_reference(fields["dens"])))


def test_c3_profiles_512_cubed_properties(cuda_device):
    """512^3 fp64 on one B200: (i) a separable known answer, (ii) fused x+z pass == single-axis passes,
    (iii) bitwise reproducibility.  u_x = a(x) + b(y) c(z)-type fields have closed-form plane statistics."""
    import torch

    from fava_b200 import device

    n = 512
    dev = cuda_device
    idx = torch.arange(n, device=dev, dtype=torch.float64)
    ax = torch.sin(2 * np.pi * (idx + 0.5) / n)
    by = torch.cos(4 * np.pi * (idx + 0.5) / n)
    rho = (1.0 + 0.25 * by).view(1, n, 1).expand(n, n, n).contiguous()  # rho(y)
    ux = (3.0 + ax.view(1, 1, n) + by.view(1, n, 1)).expand(n, n, n).contiguous()  # a(x) + b(y)
    uy = (ax.view(n, 1, 1) * torch.ones(1, n, n, device=dev, dtype=torch.float64)).contiguous()  # a(z)
    uz = torch.full((n, n, n), -2.0, device=dev, dtype=torch.float64)
    cv, lv = 1.0 / n**3, 1.0 / n
    out = device.plane_profiles(rho, ux, uy, uz, 0, cv, lv)  # planes of constant x
    # in a plane x = const: u_x - mean = b(y) - <b> = b(y) (zero mean), rho = 1 + b/4  =>  Rxx = <rho b^2> = 1/2
    maxnorm_close(out["means"][1].cpu().numpy(), (3.0 + ax).cpu().numpy(), 1e-13, "mean u_x(x)")
    maxnorm_close(out["reynolds"][0].cpu().numpy(), np.full(n, 0.5), 1e-12, "Rxx(x)")
    # u_y = a(z) varies inside the plane with zero mean: Ryy = <rho> <a^2> = 1/2, Rxy = <rho b><a> = 0
    maxnorm_close(out["reynolds"][3].cpu().numpy(), np.full(n, 0.5), 1e-12, "Ryy(x)")
    assert float(out["reynolds"][1].abs().max()) < 1e-13 and float(out["reynolds"][5].abs().max()) == 0.0
    # Favre mean of u_x: <rho u_x>/<rho> = 3 + a(x) + <(1 + b/4) b> = 3 + a(x) + 1/8
    maxnorm_close(out["favre_means"][0].cpu().numpy(), (3.125 + ax).cpu().numpy(), 1e-13, "Favre mean u_x(x)")
    (mx, px), (mz, pz) = device.plane_moments_xz(rho, ux, uy, uz)
    fx = device.moments_finalize(mx, px, cv, lv)
    for k in out:
        maxnorm_close(fx[k].cpu().numpy(), out[k].cpu().numpy(), 1e-13, f"fused xz vs axis 0: {k}")
    fz = device.moments_finalize(mz, pz, cv, lv)
    oz = device.plane_profiles(rho, ux, uy, uz, 2, cv, lv)
    for k in oz:
        maxnorm_close(fz[k].cpu().numpy(), oz[k].cpu().numpy(), 1e-13, f"fused xz vs axis 2: {k}")
    again = device.plane_profiles(rho, ux, uy, uz, 0, cv, lv)
    assert all(torch.equal(out[k], again[k]) for k in out)


def test_c4_spectrum_512_cubed_identities(cuda_device):
    """512^3 fp64 spectrum on one GPU: transverse = total - longitudinal exactly as the reference defines it, a
    band-limited solenoidal-free test field puts all energy in the expected shells, run-to-run bitwise equal."""
    import torch

    from fava_b200 import device

    n = 512
    dev = cuda_device
    idx = (torch.arange(n, device=dev, dtype=torch.float64)) / n
    x = idx.view(1, 1, n)
    y = idx.view(1, n, 1)
    z = idx.view(n, 1, 1)
    rho = torch.ones((n, n, n), device=dev, dtype=torch.float64)
    ux = (torch.cos(2 * np.pi * 7 * x) + 0.5 * torch.cos(2 * np.pi * (3 * y + 4 * z))).expand(n, n, n).contiguous()
    uy = torch.zeros((n, n, n), device=dev, dtype=torch.float64)
    uz = (0.25 * torch.sin(2 * np.pi * 12 * y)).expand(n, n, n).contiguous()
    sp = device.ke_spectrum(rho, ux, uy, uz)
    assert sp["k"].shape == (n // 2 - 1,) and np.array_equal(sp["k"], np.arange(n // 2 - 1, dtype=np.float64))
    maxnorm_close(sp["transverse"], sp["total"] - sp["longitudinal"], 1e-13, "transverse = total - longitudinal")
    # modes: |k| = 7 (amplitude 1), |k| = 5 (amplitude 1/2), |k| = 12 (amplitude 1/4); everything else is zero
    k = np.arange(-n // 2, n // 2)

    def count(m):  # lattice points of shell m (exact integer test, as the kernel does)
        k2 = (k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2)
        return int(np.sum((k2 > m * m - m) & (k2 <= m * m + m)))

    expect = np.zeros(n // 2 - 1)
    for m, amp in ((7, 1.0), (5, 0.5), (12, 0.25)):
        expect[m] = 4 * np.pi * m**2 * (2 * 0.5 * (amp / 2) ** 2) / count(m)
    maxnorm_close(sp["total"], expect, 1e-12, "band-limited total spectrum")
    again = device.ke_spectrum(rho, ux, uy, uz)
    assert all(np.array_equal(sp[q], again[q]) for q in sp)
