"""Slab-parallel Reynolds/Favre plane profiles on device tensors.

Each rank holds a z-slab [nz_local][ny][nx] of rho, ux, uy, uz (FLASH file layout).  Replaces the
MPI pattern of FLASH.reynolds_stress — per-rank partial plane sums + 4 + 6 Allreduce calls of N
doubles (_flash.py:1579-1582, :1606-1609) — by one packed [14][N] fp64 all-reduce for axis x / y
(planes span all slabs) and an all-gather for axis z (planes are slab-local).
"""

from __future__ import annotations

import torch

from fava_b200 import device, dist


def slab_profiles(rho, ux, uy, uz, axis: int, cell_volume: float, layer_volume: float, favre: bool = True,
                  gather: bool = True) -> dict[str, torch.Tensor]:
    """Profiles of the global grid formed by stacking the ranks' slabs along z.

    Pivots: for axis 0/1 a plane crosses every slab, so all ranks must share one pivot per plane:
    rank 0's (its slab holds the plane's first cell, z = 0) is broadcast.  For axis 2 every plane
    lives on one rank and the pivot stays local.
    """
    if axis in (0, 1):
        piv = device.plane_pivots(ux, uy, uz, axis)
        dist.broadcast_(piv, src=0)
        mom, _ = device.plane_moments(rho, ux, uy, uz, axis, pivots=piv)
        dist.allreduce_sum_(mom)
        return device.moments_finalize(mom, piv, cell_volume, layer_volume, favre=favre)
    out = device.plane_profiles(rho, ux, uy, uz, axis, cell_volume, layer_volume, favre=favre)
    if gather and dist.world_size() > 1:
        out = {k: dist.all_gather_cat(v, dim=1) for k, v in out.items()}
    return out
