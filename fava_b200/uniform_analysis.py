"""Uniform-grid analyses next to the spectrum: box-counting fractal dimension of an
iso-contour and velocity structure functions — the drop-ins for `FlashUniform.fractal_dimension`
(fava/mesh/FLASH/FlashUniform.py:85-227) and `FlashUniform.structure_functions` (:306-445).

Split of work:
  * fractal dimension — the field never leaves HBM: edge marking + box counts of every level are two kernels
    (csrc/fractal.cu); the host only fits the 6-12 (level, log2 count) points with the reference's formulas.
    Ranks own z-ranges aligned to the kernel's 32-plane tiles (plus one halo plane), staged straight from the
    file; the per-level counts and the tile-occupancy grid are summed over ranks.
  * structure functions — the point pairs come from NumPy's global RandomState on rank 0 with the reference's own
    sequence of draws (np.random.seed(...) therefore reproduces the reference's sample; this stream is the
    contract, not a fallback); cell lookup, the six gathers per pair, projections, powers and the sums over the
    points of each separation run on the GPU (csrc/structure.cu).  Ranks gather from their z-slabs and sum.
There is no CPU path for the field data: without the CUDA library every call raises.
"""

from __future__ import annotations

from math import log2

import numpy as np
import torch

from fava_b200 import device, dist
from fava_b200._lib import FAVA_FRACTAL_MAXLEVELS

TILE = 32  # planes per tile of fava_fractal_tiles


# ---- host helpers (pure NumPy; covered by the CPU tests) ---------------------------------------------
def box_levels(dims) -> int:
    """Number of box sizes 1, 2, 4, ... the reference counts: int(log2(min dim) + 1) (FlashUniform.py:179-184)."""
    return int(log2(min(int(v) for v in dims)) + 1)


def tile_plane_range(nz: int, rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Planes [z0, z1) a rank flags: a contiguous share of the 32-plane tiles (the last tile may be partial)."""
    tiles = (int(nz) + TILE - 1) // TILE
    t0, t1 = dist.parallel_range(tiles, rank, world)
    return min(t0 * TILE, int(nz)), min(t1 * TILE, int(nz))


def box_count_fit(counts) -> dict:
    """Filled-box counts per level -> the reference's result dict of one contour (FlashUniform.py:204-226):
    mean successive log2 ratio, and slope / R^2 / intercept of log2(count) against (levels - 1 - level)."""
    n = len(counts)
    pts = np.zeros((n, 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        for level in range(n):
            pts[level] = (n - level - 1, np.log2(int(counts[level])))
        filled = 2 ** pts[:, 1]
        avg = np.sum(np.log2(filled[:-1] / filled[1:])) / (filled.size - 1.0)
        mean, std = np.mean(pts, axis=0), np.std(pts, axis=0)
        rval = np.sum((pts[:, 0] - mean[0]) * (pts[:, 1] - mean[1])) / (np.prod(std) * n)
        slope = rval * std[1] / std[0]
        fit = np.array([slope, rval**2, mean[1] - slope * mean[0]])
    return {"average fractal dimension": avg, "slope": fit[0], "R2": fit[1], "curve": fit[2]}


def draw_point_pairs(bounds: np.ndarray, sep: float, num_points: int) -> tuple[np.ndarray, np.ndarray]:
    """One separation's pairs from np.random's global state, draw for draw like FlashUniform.py:361-395: the first
    points uniform in the domain, then azimuth and polar angle of the offset of length `sep`; the second points
    are wrapped back into the periodic domain."""
    ndim = bounds.shape[0]
    first = np.random.random((num_points, ndim)) * np.diff(bounds, axis=1).ravel() + bounds[:, 0].ravel()
    phi = 2.0 * np.pi * np.random.random(num_points)
    theta = np.arccos(2.0 * np.random.random(num_points) - 1.0)
    second = np.empty_like(first)
    second[:, 0] = first[:, 0] + sep * np.sin(theta) * np.cos(phi)
    second[:, 1] = first[:, 1] + sep * np.sin(theta) * np.sin(phi)
    second[:, 2] = first[:, 2] + sep * np.cos(theta)
    for ax in range(3):
        lo, hi = bounds[ax]
        col = second[:, ax]  # view
        while np.any(col > hi):
            col[col > hi] += lo - hi
        while np.any(col < lo):
            col[col < lo] += hi - lo
    return first, second


# ---- drivers -----------------------------------------------------------------------------------------------
def fractal_dimension(mesh, field: str, contours=0.5) -> dict:
    """{field: {str(contour): {"average fractal dimension", "slope", "R2", "curve"}}}.  The reference accepts a
    single float only (anything else raises ValueError, FlashUniform.py:87-90); lists of floats are accepted too."""
    if isinstance(contours, float):
        levels_of = [contours]
    elif isinstance(contours, (list, tuple)) and contours and all(isinstance(c, float) for c in contours):
        levels_of = list(contours)
    else:
        raise ValueError("Contours must be either a float or list of floats")
    nx, ny, nz = (int(v) for v in mesh.nCellsVec)
    if int(mesh.ndim) != 3 or nz == 1:
        raise NotImplementedError("fractal_dimension is implemented for 3-D datasets")
    nlev = box_levels((nx, ny, nz))
    edge = 2 ** (nlev - 1)
    if nx % edge or ny % edge or nz % edge or nlev > FAVA_FRACTAL_MAXLEVELS:
        # the reference indexes out of bounds (IndexError) once a box sticks out of the grid, FlashUniform.py:197-202
        raise ValueError(f"box counting needs every grid dimension to be a multiple of {edge}; got {(nx, ny, nz)}")
    key = mesh._resolve(field)
    if key is None:
        raise KeyError(f"{field} field not found in dataset {mesh.filename}")

    world = dist.world_size()
    if world == 1:
        z0, z1, zf0 = 0, nz, 0
        data = mesh.device_data(key)
        data = data.reshape(data.shape[-3:])
    else:
        z0, z1 = tile_plane_range(nz)
        zf0 = max(z0 - 1, 0)
        data = mesh._stage_plane_range(key, zf0, min(z1 + 1, nz)) if z1 > z0 else None
    dev = data.device if data is not None else torch.device("cuda", torch.cuda.current_device())
    tiles = [(v + TILE - 1) // TILE for v in (nz, ny, nx)]

    out: dict = {}
    for contour in levels_of:
        counts = torch.zeros(FAVA_FRACTAL_MAXLEVELS, dtype=torch.int64, device=dev)
        coarse = torch.zeros(tiles, dtype=torch.uint8, device=dev)
        if z1 > z0:
            device.fractal_tiles(data, contour, nz, zf0, z0, z1, counts, coarse)
        dist.allreduce_sum_(counts)
        dist.allreduce_sum_(coarse)  # tiles are owned by exactly one rank
        device.fractal_coarse(coarse, (nz, ny, nx), nlev, counts)
        out[f"{contour}"] = box_count_fit(counts[:nlev].cpu().numpy())
    return {field: out}


def structure_functions(mesh, num_seps: int = 100, num_points: int = 10000, sep_bounds=[0.0, 1.0],
                        log_scale: bool = True, anistropic: bool = False) -> dict:
    """{"transverse": {"1".."10": [num_seps]}, "longitudinal": {...}, "separations": [num_seps]}; a fresh random
    sample per order, as in the reference (the keyword is spelt `anistropic` there)."""
    if int(mesh.ndim) != 3:
        raise NotImplementedError("structure_functions is implemented for 3-D datasets")
    separations = np.geomspace(*sep_bounds, num_seps) if log_scale else np.linspace(*sep_bounds, num_seps)
    bounds = np.asarray(mesh.domain_bounds, dtype=np.float64)
    cell = np.diff(bounds, axis=1).flatten() / mesh.nCellsVec
    vel = [mesh.device_data(k) for k in ("velx", "vely", "velz")]
    vel = [v.reshape(v.shape[-3:]) for v in vel]
    nz = int(mesh.nCellsVec[2])
    zf0 = mesh._part[1] if mesh._part is not None and mesh._part[0] == "slab" else 0
    dev = vel[0].device
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    shape = (num_seps, num_points, 3)

    out: dict = {"transverse": {}, "longitudinal": {}}
    for order in range(1, 11):
        if dist.is_root():
            p1h, p2h = np.empty(shape), np.empty(shape)
            for i in range(num_seps):
                p1h[i], p2h[i] = draw_point_pairs(bounds, separations[i], num_points)
            p1, p2 = torch.from_numpy(p1h).to(dev), torch.from_numpy(p2h).to(dev)
        else:
            p1, p2 = (torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(2))
        dist.broadcast_(p1)
        dist.broadcast_(p2)
        v1 = device.sf_gather(p1.view(-1, 3), *vel, nz, zf0, bounds[:, 0], cell, err)
        v2 = device.sf_gather(p2.view(-1, 3), *vel, nz, zf0, bounds[:, 0], cell, err)
        dist.allreduce_sum_(v1)  # exactly one rank holds each cell; the others add 0.0
        dist.allreduce_sum_(v2)
        sums = device.sf_moments(p1, p2, v1, v2, num_seps, num_points, order, anistropic).cpu().numpy()
        if int(err.item()):
            raise IndexError("structure_functions: a sample point lies on the upper domain face (index out of bounds "
                             "for the grid), as in the reference")
        out["longitudinal"][f"{order}"] = sums[0].copy()
        out["transverse"][f"{order}"] = sums[1].copy()
        out["separations"] = separations
    return out
