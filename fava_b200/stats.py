"""Slab-parallel Reynolds/Favre plane profiles on device tensors.

Each rank holds a z-slab [nz_local][ny][nx] of rho, ux, uy, uz (FLASH file layout).  Replaces the
MPI pattern of FLASH.reynolds_stress — per-rank partial plane sums + 4 + 6 Allreduce calls of N
doubles (_flash.py:1579-1582, :1606-1609) — by one packed [14][N] fp64 all-reduce for axis x / y
(planes span all slabs) and an all-gather for axis z (planes are slab-local).
"""

from __future__ import annotations

import torch

from fava_b200 import device, dist


def slab_moments_local(rho, ux, uy, uz, axis: int):
    """Kernel-only half of `slab_profiles`: this rank's pivoted plane moments (no collective)."""
    return device.plane_moments(rho, ux, uy, uz, axis)


def slab_profiles_finish(mom, piv, axis: int, cell_volume: float, layer_volume: float, favre: bool = True,
                         gather: bool = True) -> dict[str, torch.Tensor]:
    """Collective half: for axis 0/1 a plane crosses every slab, so the ranks first agree on ONE pivot per
    plane (rank 0's — its slab holds the plane's first cell, z = 0), re-express their moments about it
    (exact algebra) and add them with one packed [14][N] all-reduce; for axis 2 every plane lives on one
    rank and the profiles are simply concatenated."""
    if axis in (0, 1):
        if dist.world_size() > 1:
            piv = agree_pivots(mom, piv)
            dist.allreduce_sum_(mom)
        return device.moments_finalize(mom, piv, cell_volume, layer_volume, favre=favre)
    out = device.moments_finalize(mom, piv, cell_volume, layer_volume, favre=favre)
    if gather and dist.world_size() > 1:
        out = {k: dist.all_gather_cat(v, dim=1) for k, v in out.items()}
    return out


def slab_profiles(rho, ux, uy, uz, axis: int, cell_volume: float, layer_volume: float, favre: bool = True,
                  gather: bool = True) -> dict[str, torch.Tensor]:
    """Profiles of the global grid formed by stacking the ranks' slabs along z."""
    mom, piv = slab_moments_local(rho, ux, uy, uz, axis)
    return slab_profiles_finish(mom, piv, axis, cell_volume, layer_volume, favre=favre, gather=gather)


def agree_pivots(mom: torch.Tensor, piv: torch.Tensor) -> torch.Tensor:
    """Block-range ranks pick their pivots from their own first block plane per bin; before the
    moments can be summed across ranks they must refer to ONE pivot per bin.  Every rank adopts the
    pivot of the lowest rank that actually holds data for the bin (weight row > 0) and re-expresses
    its moments about it (exact algebra, fava_moments_repivot)."""
    if dist.world_size() == 1:
        return piv
    has = (mom[13] > 0).to(torch.float64)
    packed = torch.cat([piv, has[None, :]], dim=0)
    allp = dist.all_gather_rows(packed)
    common = torch.zeros_like(piv)
    chosen = torch.zeros_like(has)
    for part in allp:  # rank order
        take = (chosen == 0) & (part[3] > 0)
        common = torch.where(take[None, :], part[:3], common)
        chosen = torch.where(take, torch.ones_like(chosen), chosen)
    device.moments_repivot(mom, piv, common)
    return common


def block_profiles(rho, ux, uy, uz, axis: int, table, nbins: int, layer_volume: float,
                   favre: bool = True) -> dict[str, torch.Tensor]:
    """Profiles of a block dataset whose leaves are sharded over ranks as contiguous block ranges
    (the reference's MPI decomposition, _flash.py:166-208, :803-822)."""
    mom, piv = device.plane_moments_blocks(rho, ux, uy, uz, axis, table, nbins)
    if dist.world_size() > 1:
        piv = agree_pivots(mom, piv)
        dist.allreduce_sum_(mom)
    return device.moments_finalize(mom, piv, 1.0, layer_volume, favre=favre)


def slab_plane_sum(field, axis: int, cell_volume: float) -> torch.Tensor:
    """vf-weighted plane integral of one field held as z-slabs (reference slice_integral, _flash.py:1451-1504)."""
    out = device.plane_sum(field, axis) * cell_volume
    if axis in (0, 1):
        return dist.allreduce_sum_(out)
    return dist.all_gather_cat(out, dim=0)


def block_plane_sum(blocks, axis: int, table, nbins: int) -> torch.Tensor:
    out = device.plane_sum_blocks(blocks, axis, table, nbins)
    return dist.allreduce_sum_(out)


def slab_step(rho, ux, uy, uz, n: int, cell_volume: float, layer_volume: float, axes=(0, 1, 2), spectrum: bool = True,
              favre: bool = True) -> dict:
    """One full statistics pass over a snapshot held as z-slabs: plane profiles along `axes` plus the
    kinetic-energy spectrum, scheduled so that the moment kernels (HBM-bound, no communication) run while
    the spectrum's slab exchange (NVLink-bound) is in flight; the profile collectives follow afterwards so
    that they never queue behind the exchange fences.  Returns {axis: profile dict, ..., "spectrum": dict}."""
    from fava_b200 import spectrum as spec

    out, pending = {}, {}

    def moments_xz():  # x and z bins from ONE pass over the slab
        pending[0], pending[2] = device.plane_moments_xz(rho, ux, uy, uz)

    def moments_of(ax):
        def run():
            pending[ax] = slab_moments_local(rho, ux, uy, uz, ax)

        return run

    def moments_xyz():  # x, y and z bins from ONE pass over the slab
        (pending[0], pending[1], pending[2]) = device.plane_moments_xyz(rho, ux, uy, uz)

    pieces = []
    todo = list(axes)
    if set(todo) >= {0, 1, 2} and device.plane_moments_xyz_supported(rho.shape):
        pieces.append(moments_xyz)
        todo = []
    elif 0 in todo and 2 in todo:
        pieces.append(moments_xz)
        todo = [ax for ax in todo if ax == 1]
    pieces += [moments_of(ax) for ax in todo]

    def local_moments():
        for piece in pieces:
            piece()

    def finish_profiles():
        for ax in axes:
            mom, piv = pending[ax]
            out[ax] = slab_profiles_finish(mom, piv, ax, cell_volume, layer_volume, favre=favre, gather=False)

    if spectrum:
        out["spectrum"] = spec.slab_ke_spectrum(rho, ux, uy, uz, n, overlap=pieces, epilogue=finish_profiles)
    else:
        local_moments()
        finish_profiles()
    return out


def host_step(host, n: int, cell_volume: float, layer_volume: float, axes=(0, 1, 2), spectrum: bool = True,
              favre: bool = True, chunk_planes: int = 64, stage=None) -> dict:
    """`slab_step` for a snapshot that still lives in HOST memory (pinned tensors rho, ux, uy, uz of this rank's
    z-slab [nz_local][n][n]): the slab is copied to HBM in chunks of `chunk_planes` planes on a side stream and
    every chunk is consumed as soon as it has landed — plane moments accumulated about the pivots of the first
    chunk, weighting + x and y transforms written into the spectral buffers — so that all the HBM-bound work except
    the z transforms and the binning hides behind the PCIe copy.  Same result dict as `slab_step`."""
    from fava_b200 import spectrum as spec

    nzl = int(host[0].shape[0])
    dev = torch.device("cuda", torch.cuda.current_device())
    if stage is None:
        stage = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host]
    cur = torch.cuda.current_stream(dev)
    copy_stream = _copy_stream(dev)
    copy_stream.wait_stream(cur)  # earlier users of `stage` on the calling stream
    chunks = [(a, min(a + chunk_planes, nzl)) for a in range(0, nzl, chunk_planes)]
    landed = []
    with torch.cuda.stream(copy_stream):
        for a, b in chunks:
            for h, s in zip(host, stage):
                s[a:b].copy_(h[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            landed.append(ev)

    w = spec.spectral_buffers(n, nzl, dev) if spectrum else None
    plan = spec._plan(n, dist.rank(), dist.world_size(), dev) if spectrum and dist.world_size() > 1 else None
    plane_bytes = device.spectral_bytes(n, 1) if spectrum else 0  # one z-plane of a spectral buffer: complex [n][pitch]
    mom, piv = {}, {}
    for (a, b), ev in zip(chunks, landed):
        cur.wait_event(ev)
        part = [s[a:b] for s in stage]
        for ax in axes:
            if ax in (0, 1):  # the planes cross every chunk: accumulate about the first chunk's pivots
                if a == 0:
                    mom[ax], piv[ax] = device.plane_moments(*part, ax)
                else:
                    device.plane_moments(*part, ax, pivots=piv[ax], out=mom[ax], accumulate=True)
            else:  # z planes are chunk-local
                m, pv = device.plane_moments(*part, ax)
                if a == 0:
                    mom[ax] = torch.empty((m.shape[0], nzl), dtype=m.dtype, device=dev)
                    piv[ax] = torch.empty((pv.shape[0], nzl), dtype=pv.dtype, device=dev)
                mom[ax][:, a:b] = m
                piv[ax][:, a:b] = pv
        if spectrum:
            if plan is None or not plan.native:  # (cuFFT path on several ranks: the rows are exchanged after the last chunk)
                device.ke_transform_xy(*part, *[p + a * plane_bytes for p in w])
            else:  # several ranks: x pass of the chunk, then its y pass scatters the rows to their owners
                device.ke_transform_x(*part, *[p + a * plane_bytes for p in w])
                for c in range(3):
                    spec.exchange(plan, c, z_offset=a, nz_chunk=b - a)
    out = {}

    def finish_profiles():
        for ax in axes:
            out[ax] = slab_profiles_finish(mom[ax], piv[ax], ax, cell_volume, layer_volume, favre=favre, gather=False)

    if spectrum:
        out["spectrum"] = spec.spectrum_from_transformed_slabs(n, dev, epilogue=finish_profiles)
    else:
        finish_profiles()
    return out


_copy_streams: dict = {}


def _copy_stream(dev) -> torch.cuda.Stream:
    key = str(dev)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=dev)
    return _copy_streams[key]
