"""Closed-form known-answer snapshots for any grid size and any z-slab split.

Two synthetic fields whose plane statistics / shell spectrum follow from a few lines of algebra, so that a run at
a size no CPU oracle can reach (1024^3 needs ~250 GB for the reference) and on any number of ranks
can still be checked against numbers that do not come from this code base:

* profile case  rho = 1 + b(y)/4,  u_x = 3 + a(x) + b(y),  u_y = a(z),  u_z = -2  with
  a(t) = sin(2 pi (t + 1/2)/n), b(t) = cos(4 pi (t + 1/2)/n): every plane mean and every Reynolds / Favre stress
  (reference definition, fava/mesh/FLASH/_flash.py:1564-1609) is a low-order trigonometric average;
* spectrum case  rho = 1,  u_x = cos(2 pi 7 x) + cos(2 pi (3 y + 4 z))/2,  u_y = 0,  u_z = sin(2 pi 12 y)/4: three
  Fourier modes of moduli 7, 5 and 12, so `total` of FlashUniform.kinetic_energy_spectra
  (fava/mesh/FLASH/FlashUniform.py:261, :273-302) is non-zero in exactly three shells with values fixed by the
  lattice-point count of each shell.

Used by bench.py (`parity_check` in the JSON line, every N) and by the GPU tests; nothing here touches `oracle/`.
"""

from __future__ import annotations

import numpy as np
import torch

STRESS_ROWS = ("Rxx", "Rxy", "Rxz", "Ryy", "Ryz", "Rzz")


def _ab(n: int, dev):
    idx = torch.arange(n, device=dev, dtype=torch.float64)
    return torch.sin(2 * np.pi * (idx + 0.5) / n), torch.cos(4 * np.pi * (idx + 0.5) / n)


def fill_profile_case(fields, n: int, z0: int) -> None:
    """Overwrite the slab tensors rho, ux, uy, uz ([nz_local][n][n], planes z0...) in place."""
    rho, ux, uy, uz = fields
    nz = int(rho.shape[0])
    a, b = _ab(n, rho.device)
    rho.copy_((1.0 + 0.25 * b).view(1, n, 1).expand(nz, n, n))
    ux.copy_((3.0 + a.view(1, 1, n) + b.view(1, n, 1)).expand(nz, n, n))
    uy.copy_(a[z0:z0 + nz].view(nz, 1, 1).expand(nz, n, n))
    uz.fill_(-2.0)


def expected_profiles(n: int, axis: int) -> dict[str, np.ndarray]:
    """Global profiles along `axis` of the profile case: means [4][n], reynolds [6][n], favre_means [3][n], favre [6][n]."""
    t = (np.arange(n) + 0.5) / n
    a, b = np.sin(2 * np.pi * t), np.cos(4 * np.pi * t)
    one, zero = np.ones(n), np.zeros(n)
    means = np.zeros((4, n))
    rey = np.zeros((6, n))
    fmeans = np.zeros((3, n))
    fav = np.zeros((6, n))
    means[3] = fmeans[2] = -2.0
    if axis == 0:  # planes x = const: u_x' = b(y), u_y' = a(z)
        means[0], means[1] = one, 3.0 + a
        rey[0], rey[3] = 0.5 * one, 0.5 * one  # <rho b^2> = 1/2 + <b^3>/4,  <rho><a^2>
        fmeans[0] = 3.0 + a + 0.125  # + <rho b>/<rho>
        fav[0] = 0.5 - 1.0 / 64.0  # <rho b^2> - <rho b>^2/<rho>
        fav[3] = 0.5 * one
    elif axis == 1:  # planes y = const: rho constant in the plane, u_x' = a(x), u_y' = a(z)
        r = 1.0 + 0.25 * b
        means[0], means[1] = r, 3.0 + b
        rey[0], rey[3] = 0.5 * r, 0.5 * r
        fmeans[0] = 3.0 + b
        fav[0], fav[3] = 0.5 * r, 0.5 * r
    else:  # planes z = const: u_x' = a(x) + b(y), u_y constant in the plane
        means[0], means[1], means[2] = one, 3.0 * one, a
        rey[0] = one  # <rho><a^2> + <rho b^2>
        fmeans[0], fmeans[1] = 3.125 * one, a
        fav[0] = 1.0 - 1.0 / 64.0
    del zero
    return {"means": means, "reynolds": rey, "favre_means": fmeans, "favre": fav}


def fill_spectrum_case(fields, n: int, z0: int) -> None:
    rho, ux, uy, uz = fields
    nz = int(rho.shape[0])
    dev = rho.device
    idx = torch.arange(n, device=dev, dtype=torch.float64) / n
    x, y = idx.view(1, 1, n), idx.view(1, n, 1)
    z = idx[z0:z0 + nz].view(nz, 1, 1)
    rho.fill_(1.0)
    ux.copy_((torch.cos(2 * np.pi * 7 * x) + 0.5 * torch.cos(2 * np.pi * (3 * y + 4 * z))).expand(nz, n, n))
    uy.zero_()
    uz.copy_((0.25 * torch.sin(2 * np.pi * 12 * y)).expand(nz, n, n))


def shell_count(m: int) -> int:
    """Lattice points k in Z^3 with m - 1/2 < |k| < m + 1/2 (no ties: |k|^2 is an integer)."""
    k = np.arange(-m - 1, m + 2)
    k2 = k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2
    return int(np.sum((k2 > m * m - m) & (k2 <= m * m + m)))


def expected_spectrum_total(n: int) -> np.ndarray:
    """`total` of the spectrum case: shell mean of |u^|^2 / 2 times 4 pi k^2 (FlashUniform.py:288-297)."""
    if n // 2 - 1 <= 12:
        raise ValueError("the spectrum case needs n >= 28")
    expect = np.zeros(n // 2 - 1)
    for m, amp in ((7, 1.0), (5, 0.5), (12, 0.25)):
        expect[m] = 4 * np.pi * m**2 * (2 * 0.5 * (amp / 2) ** 2) / shell_count(m)
    return expect


def _rel(got: np.ndarray, want: np.ndarray) -> float:
    scale = max(float(np.max(np.abs(want))), 1e-300)
    return float(np.max(np.abs(np.asarray(got) - want))) / scale


def profile_errors(result: dict, n: int, axes, z0: int, nz: int) -> dict[str, float]:
    """max-norm relative error per (axis, array) of a `slab_step` / `host_step` result on the profile case.  Axis-2
    profiles of a slab step hold this rank's planes only (gather=False); identically-zero rows are compared absolutely."""
    errs = {}
    for ax in axes:
        want = expected_profiles(n, ax)
        for key, w in want.items():
            got = result[ax][key].detach().cpu().numpy()
            if ax == 2 and got.shape[1] != n:
                w = w[:, z0:z0 + nz]
            for row in range(w.shape[0]):
                scale = float(np.max(np.abs(want[key][row])))
                e = float(np.max(np.abs(got[row] - w[row])))
                errs[f"axis{ax}.{key}[{row}]"] = e / scale if scale > 0 else e
    return errs


def spectrum_errors(spec: dict, n: int) -> dict[str, float]:
    want = expected_spectrum_total(n)
    return {
        "spectrum.k": float(np.max(np.abs(spec["k"] - np.arange(n // 2 - 1)))),
        "spectrum.total": _rel(spec["total"], want),
        "spectrum.transverse=total-longitudinal": _rel(spec["transverse"], spec["total"] - spec["longitudinal"]),
    }
