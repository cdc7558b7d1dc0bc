"""Slab-decomposed kinetic-energy spectrum (BASELINE configs[3]: 1024^3 at 1/2/4/8 B200).

Rank r holds z-planes [r*nzl, (r+1)*nzl) of rho, ux, uy, uz.  Per velocity component:
  weight (K4) -> in-place 2-D FFT over (y,x) of the local planes (cuFFT D2Z) -> slab->pencil exchange
  (K5: ONE kernel gathers the ky rows each destination owns and stores them straight into that rank's
  receive buffer over NVLink peer mappings) -> 1-D FFT along z (cuFFT Z2Z) -> shell binning (K6) ->
  one all-reduce of the [3][N/2-1] shell sums.
Spectral space is distributed over ky in +-ky symmetric sets, so the transposed operand
u^(kz, +-ky, kx) the reference's `.T` projection needs (FlashUniform.py:281) is always rank-local.
"""

from __future__ import annotations

import numpy as np
import torch

from fava_b200 import device, dist


def ky_ownership(n: int, nranks: int) -> np.ndarray:
    """[nranks][nyl] global ky indices owned by each rank (-1 = padding).  Rank r owns the wavenumbers
    |ky| in [r*h, (r+1)*h), h = N/(2P), as index pairs (j, N-j); the Nyquist row j = N/2 is owned by
    nobody (it lies beyond the last bin edge).  nyl = 2h."""
    if n % (2 * nranks):
        raise ValueError(f"grid size {n} must be divisible by 2 x {nranks} ranks")
    h = n // (2 * nranks)
    own = -np.ones((nranks, 2 * h), dtype=np.int32)
    for r in range(nranks):
        pos = np.arange(r * h, (r + 1) * h)
        neg = (n - pos) % n
        rows = list(pos) + [j for j in neg[::-1] if j not in pos]
        own[r, : len(rows)] = rows
    return own


def slab_ke_spectrum(rho, ux, uy, uz, n: int) -> dict[str, np.ndarray]:
    """Spectrum of the global N^3 grid formed by the ranks' z-slabs; every rank returns the full dict."""
    world, rank = dist.world_size(), dist.rank()
    if world == 1:
        return device.ke_spectrum(rho, ux, uy, uz)
    return device.ke_spectrum_slab(rho, ux, uy, uz, n, rank, world)
