"""CPU oracle for the FAVA grid-statistics hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the reference's algorithms (ebrooker/FAVA, paths relative to the reference
root).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; nothing under `fava_b200/` does, and the product path has no CPU
fallback.

Parity pin: the reference ships no golden vectors or known-answer tests for this path (its tests
only touch base-class names, tests/test_model.py:20-24).  The oracle is therefore pinned against
outputs of the *unmodified reference itself*, executed in the build container on deterministic
synthetic FLASH files through import shims (oracle/ref_harness.py); the vectors and the script
that made them are committed under tests/golden/ (make_golden.py) and checked by
tests/test_oracle_golden.py.

Array layout: functions here take arrays in the reference's IN-MEMORY layout, i.e. what
`FLASH._read_variable_data` produces (_flash.py:306-341): float64, axes -1 and -3 of the file
dataset swapped, so a block dataset is [blk, i(x), j(y), k(z)] and a uniform one [i, j, k].
`load_like_reference` converts a file-layout array ([...,z,y,x]).
"""

from __future__ import annotations

import itertools

import numpy as np

MESH_MDIM = 3  # fava/mesh/FLASH/_util.py:16


# ------------------------------------------------------------------------------------------------
# A1 — loader (fava/mesh/FLASH/_flash.py:331-335)
# ------------------------------------------------------------------------------------------------
def load_like_reference(file_array: np.ndarray) -> np.ndarray:
    """dataset[()] -> float64 -> swap axes -1/-3 -> C-contiguous, as _flash.py:332-334."""
    return np.ascontiguousarray(np.swapaxes(np.asarray(file_array).astype(np.float64), -1, -3))


# ------------------------------------------------------------------------------------------------
# A4' — geometry helpers (_flash.py:591-617, :914-953)
# ------------------------------------------------------------------------------------------------
class MeshGeom:
    """The O(nblocks) metadata `reynolds_stress` / `from_amr` consume (set by FLASH.load, :106-163)."""

    def __init__(self, ncells_vec, nblks_vec, domain_bounds, block_bounds, refine_level, node_type, ndim=3):
        self.nCellsVec = np.asarray(ncells_vec, dtype=np.int32)  # _flash.py:406-407
        self.nBlksVec = np.asarray(nblks_vec, dtype=np.int32)  # :410-411
        self.domain_bounds = np.asarray(domain_bounds, dtype=np.float64)  # :396-399
        self.block_bounds = np.asarray(block_bounds)  # file dtype kept (f32 for plt), :350-354
        self.refine_level = np.asarray(refine_level, dtype=np.int64)  # :294-296
        self.node_type = np.asarray(node_type, dtype=np.int64)  # :290-292
        self.ndim = int(ndim)
        self.nblocks = int(self.refine_level.shape[0])

    @property
    def refine_level_max(self):  # :587-589
        return self.refine_level.max()

    @property
    def domain_volume(self):  # :591-602 (cartesian only)
        return np.prod(np.diff(self.domain_bounds))

    def min_delta(self, axis):  # get_minimum_deltas :914-917
        d = self.domain_bounds
        return (d[axis, 1] - d[axis, 0]) / (self.nCellsVec[axis] * self.nBlksVec[axis] * 2 ** (self.refine_level_max - 1))

    def delta_from_level(self, axis, level):  # get_delta_from_refine_level :930-933
        d = self.domain_bounds
        return (d[axis, 1] - d[axis, 0]) / (self.nCellsVec[axis] * self.nBlksVec[axis] * 2 ** (level - 1))

    def cell_volume_from_level(self, level=1):  # get_cell_volume_from_refinement :946-953
        cells = self.nCellsVec[0] * self.nBlksVec[0] * 2 ** (level - 1)
        if self.ndim > 1:
            cells = cells * (self.nCellsVec[1] * self.nBlksVec[1] * 2 ** (level - 1))
        if self.ndim > 2:
            cells = cells * (self.nCellsVec[2] * self.nBlksVec[2] * 2 ** (level - 1))
        return self.domain_volume / float(cells)

    def leaf_blocks(self, lb=0, ub=None):  # get_blocklist("LEAF") :803-817 for rank range [lb,ub)
        ub = self.nblocks if ub is None else ub
        return (lb + np.argwhere(self.node_type[lb:ub] == 1).flatten()).astype(np.int64)


def uniform_geom(shape_xyz, domain_bounds=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), bbox_dtype=np.float32) -> MeshGeom:
    """Geometry of a single-block 'uniform' file: one leaf at level 1 covering the whole domain."""
    db = np.asarray(domain_bounds, dtype=np.float64)
    return MeshGeom(shape_xyz, (1, 1, 1), db, db.astype(bbox_dtype)[None, ...], [1], [1])


# ------------------------------------------------------------------------------------------------
# A3 + A4 — FLASH.reynolds_stress (_flash.py:1506-1611)
# ------------------------------------------------------------------------------------------------
def reynolds_stress(geom: MeshGeom, data: dict, axis: int = 0, blk_range=None):
    """Restatement of _flash.py:1506-1611 for data[key] = float64[blk, i, j, k].

    For axis 0 this is the reference's arithmetic step for step.  For axis 1/2 the reference still
    reduces over array axes (1,2) and indexes array axis 0 (:1570, :1599-1601 — it returns the
    x-profile labelled with y/z coordinates); the oracle implements the documented meaning instead:
    the same arithmetic with the roles of the array axes rotated, which equals the reference's
    raxis=0 result on the axis-permuted file (tests/golden pins that equivalence).

    Returns (radius[N+1], stress{Rxx..Rzz}[N], means{dens,velx,vely,velz}[N]).
    """
    ndim = geom.ndim
    lrefcells = 2 ** (geom.refine_level_max - 1)  # :1508
    dims = [int(nb * bl * lrefcells) for nb, bl in zip(geom.nCellsVec[:ndim], geom.nBlksVec[:ndim])]  # :1509-1511
    if axis not in (0, 1, 2):
        raise ValueError(f"Do not recognize AXIS enumeration {axis}")  # :1539-1540
    min_deltas = np.array([geom.min_delta(i) for i in range(ndim)], dtype=np.float64)  # :1515-1517
    axes = "xyz"[:ndim]
    db = geom.domain_bounds
    others = [a for a in range(3) if a != axis]
    layer_volume = (db[others[0], 1] - db[others[0], 0]) * (db[others[1], 1] - db[others[1], 0])  # :1526-1538
    rmin, rmax = db[axis, 0], db[axis, 1]
    nrb = int(geom.nCellsVec[axis])
    layer_volume = layer_volume * min_deltas[axis]  # :1542
    n = dims[axis]
    radius = np.linspace(rmin, rmax, n + 1)  # :1545

    stress = {}
    means = {"dens": np.zeros(n)}  # :1547-1554
    for i in range(ndim):
        means[f"vel{axes[i]}"] = np.zeros(n)
        for j in range(i, ndim):
            stress[f"R{axes[i]}{axes[j]}"] = np.zeros(n)

    lb, ub = (0, geom.nblocks) if blk_range is None else blk_range
    blocklist = geom.leaf_blocks(lb, ub)  # :1556
    mapping = np.zeros((blocklist.size, nrb, 2), dtype=np.int64)  # :1557
    cell_vols = np.array([geom.cell_volume_from_level(geom.refine_level[b]) for b in blocklist], dtype=np.float64)
    vol_fracs = cell_vols * (min_deltas[axis] / geom.delta_from_level(axis, geom.refine_level[blocklist]))  # :1559-1562

    # plane of constant index along `axis`: move that array axis first (a view; for axis 0 a no-op)
    def planes(arr_blk):
        return np.moveaxis(arr_blk, axis, 0)

    for l, blk in enumerate(blocklist):  # pass 1, :1564-1577
        lref_n = int(2 ** (geom.refine_level_max - 1) / 2 ** (geom.refine_level[blk] - 1))  # :1565
        lo = geom.block_bounds[blk, axis, 0]  # :1566
        ilo = int(np.argmin(np.abs(radius[:-1] - lo)))  # :1567
        _means = {key: np.einsum("ijk->i", planes(data[key][blk, ...])) * vol_fracs[l] for key in means}  # :1569-1571
        for i in range(nrb):
            jlo = ilo + i * lref_n
            jhi = ilo + (i + 1) * lref_n
            mapping[l, i, :] = [jlo, jhi]
            for key in means:
                means[key][jlo:jhi] += _means[key][i]  # :1576-1577

    for key in means:  # :1579-1582 (Allreduce over ranks is the identity for one rank)
        means[key] = means[key] / layer_volume

    for l, blk in enumerate(blocklist):  # pass 2, :1584-1604
        dens = planes(data["dens"][blk, ...])
        for i in range(ndim):
            vi = f"vel{axes[i]}"
            ui = planes(data[vi][blk, ...])
            for j in range(i, ndim):
                vj = f"vel{axes[j]}"
                uj = planes(data[vj][blk, ...])
                acc = stress[f"R{axes[i]}{axes[j]}"]
                for rk in range(nrb):
                    for ii in range(mapping[l, rk, 0], mapping[l, rk, 1]):
                        acc[ii] += (
                            np.sum(dens[rk, ...] * (ui[rk, ...] - means[vi][ii]) * (uj[rk, ...] - means[vj][ii]))
                            * vol_fracs[l]
                        )  # :1597-1604

    for key in stress:  # :1606-1609
        stress[key] = stress[key] / layer_volume
    return radius, stress, means


def favre_stress(geom: MeshGeom, data: dict, axis: int = 0):
    """Favre statistics — NOT in the reference; two-pass NumPy definition used as oracle:
    u~_i = <rho u_i>/<rho>,  F_ij = <rho (u_i-u~_i)(u_j-u~_j)>, same block/volume weighting and
    bin mapping as `reynolds_stress` above.  Returns (favre_means{velx..}, favre{Rxx..})."""
    ndim = geom.ndim
    lrefcells = 2 ** (geom.refine_level_max - 1)
    dims = [int(nb * bl * lrefcells) for nb, bl in zip(geom.nCellsVec[:ndim], geom.nBlksVec[:ndim])]
    min_deltas = np.array([geom.min_delta(i) for i in range(ndim)], dtype=np.float64)
    db = geom.domain_bounds
    others = [a for a in range(3) if a != axis]
    layer_volume = (db[others[0], 1] - db[others[0], 0]) * (db[others[1], 1] - db[others[1], 0]) * min_deltas[axis]
    n = dims[axis]
    radius = np.linspace(db[axis, 0], db[axis, 1], n + 1)
    nrb = int(geom.nCellsVec[axis])
    blocklist = geom.leaf_blocks()
    cell_vols = np.array([geom.cell_volume_from_level(geom.refine_level[b]) for b in blocklist], dtype=np.float64)
    vol_fracs = cell_vols * (min_deltas[axis] / geom.delta_from_level(axis, geom.refine_level[blocklist]))
    vel = ["velx", "vely", "velz"][:ndim]
    rho_sum = np.zeros(n)
    rhou = {v: np.zeros(n) for v in vel}
    spans = []
    for l, blk in enumerate(blocklist):
        s = int(2 ** (geom.refine_level_max - 1) / 2 ** (geom.refine_level[blk] - 1))
        ilo = int(np.argmin(np.abs(radius[:-1] - geom.block_bounds[blk, axis, 0])))
        spans.append((ilo, s))
        rho = np.moveaxis(data["dens"][blk, ...], axis, 0)
        r_pl = rho.sum(axis=(1, 2)) * vol_fracs[l]
        for i in range(nrb):
            rho_sum[ilo + i * s : ilo + (i + 1) * s] += r_pl[i]
        for v in vel:
            ru = (rho * np.moveaxis(data[v][blk, ...], axis, 0)).sum(axis=(1, 2)) * vol_fracs[l]
            for i in range(nrb):
                rhou[v][ilo + i * s : ilo + (i + 1) * s] += ru[i]
    fmeans = {v: rhou[v] / rho_sum for v in vel}
    favre = {}
    axes = "xyz"[:ndim]
    for a in range(ndim):
        for b in range(a, ndim):
            acc = np.zeros(n)
            for l, blk in enumerate(blocklist):
                ilo, s = spans[l]
                rho = np.moveaxis(data["dens"][blk, ...], axis, 0)
                ua = np.moveaxis(data[vel[a]][blk, ...], axis, 0)
                ub = np.moveaxis(data[vel[b]][blk, ...], axis, 0)
                for rk in range(nrb):
                    for ii in range(ilo + rk * s, ilo + (rk + 1) * s):
                        acc[ii] += np.sum(rho[rk] * (ua[rk] - fmeans[vel[a]][ii]) * (ub[rk] - fmeans[vel[b]][ii])) * vol_fracs[l]
            favre[f"R{axes[a]}{axes[b]}"] = acc / layer_volume
    return fmeans, favre


# ------------------------------------------------------------------------------------------------
# A6 + A7 — FlashUniform.kinetic_energy_spectra (fava/mesh/FLASH/FlashUniform.py:229-304)
# ------------------------------------------------------------------------------------------------
def _binned_mean(x: np.ndarray, values: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """scipy.stats.binned_statistic(x, values, bins=edges, statistic='mean') (scipy 1.15 pinned by the
    reference, uv.lock:460-461): bin i holds edges[i] <= x < edges[i+1], the last bin also includes
    its right edge, points outside are dropped, an empty bin gives NaN."""
    nb = edges.size - 1
    idx = np.digitize(x, edges)  # 1..nb inside, 0 / nb+1 outside
    idx[x == edges[-1]] = nb
    ok = (idx >= 1) & (idx <= nb)
    cnt = np.bincount(idx[ok] - 1, minlength=nb).astype(np.float64)
    sums = np.bincount(idx[ok] - 1, weights=values[ok], minlength=nb)
    with np.errstate(invalid="ignore", divide="ignore"):
        return sums / cnt


def kinetic_energy_spectra(data: dict, ncells_vec, ndim: int = 3, use_scipy: bool = True):
    """Restatement of FlashUniform.py:229-304 for data[key] = float64[i, j, k] (reference layout).

    Quirk kept for parity: the longitudinal projection uses `ffts[n].T` (all axes reversed,
    FlashUniform.py:281), which only works on cubic grids.
    """
    velocity = ["velx", "vely", "velz"][:ndim]  # :240
    k_num = np.asarray(ncells_vec, dtype=np.int32)[:ndim]  # :242
    k_start = -k_num // 2  # :244
    k_end = -k_start - 1  # :245
    k = np.array(np.meshgrid(*(np.linspace(ks, ke, n) for ks, ke, n in zip(k_start, k_end, k_num)), indexing="ij"))  # :248-253
    k_abs = np.abs(k) if ndim == 1 else np.sqrt((k**2).sum(axis=0))  # :256-259
    bins = np.arange(np.max(k_num) // 2) - 0.5  # :261

    dens = np.sqrt(data["dens"])  # :266
    ffts = []
    for comp in velocity:  # :267-270
        f = np.fft.fftn(dens * data[comp], norm="forward")
        ffts.append(np.fft.fftshift(f))
    ffts = np.array(ffts)  # :271

    power = {"total": 0.5 * (np.abs(ffts) ** 2).sum(axis=0)}  # :273
    lon = np.zeros(k_num, dtype=np.complex128)  # :275
    if ndim == 1:
        lon = lon + k * ffts[0, ...]  # :277
    else:
        for n in range(ndim):
            lon += k[n] * ffts[n, ...].T  # :281
    power["longitudinal"] = np.abs(lon / np.maximum(k_abs, 1e-99)) ** 2  # :283
    power["transverse"] = power["total"] - power["longitudinal"]  # :284

    spectral = {}
    for key, val in power.items():  # :287-293
        if use_scipy:
            from scipy.stats import binned_statistic

            res = binned_statistic(k_abs.flatten(), val.flatten(), bins=bins, statistic="mean")
            stat, edges = res.statistic, res.bin_edges
        else:
            stat, edges = _binned_mean(k_abs.flatten(), val.flatten(), bins), bins
        if "k" not in spectral:
            spectral["k"] = edges[:-1] + 0.5
        spectral[key] = stat
    factor = spectral["k"] ** (ndim - 1)  # :295
    if ndim > 1:
        factor = factor * (2 * np.pi * (ndim - 1))  # :297
    for key in spectral:
        if key != "k":
            spectral[key] = spectral[key] * factor  # :299-302
    return spectral


# ------------------------------------------------------------------------------------------------
# A8 — from_amr index construction (_flash.py:963-1022, :1157-1199)
# ------------------------------------------------------------------------------------------------
class AmrPlan:
    """Integer tables from_amr derives before touching field data."""

    def __init__(self):
        self.subdomain_flag = False
        self.outside = False
        self.ref_lev_max = 0
        self.grid_delta = None  # float64 (3,1)
        self.local_BCIDs = None  # int32 (nblocks,3,2)
        self.subdomain_BCIDs = None  # int32 (3,2)
        self.leaf_IDs = []
        self.total_cells = None  # int32 (3,)
        self.refdom_bound_box = None


def from_amr_plan(geom: MeshGeom, subdomain_coords, refine_level: int = -1) -> AmrPlan:
    p = AmrPlan()
    ndim = geom.ndim
    p.subdomain_flag = any(0 not in sdc for sdc in subdomain_coords)  # :965
    db = geom.domain_bounds
    if p.subdomain_flag:  # :967-977 (silent return)
        for a in range(ndim):
            if subdomain_coords[a, 0] < db[a, 0] or db[a, 1] < subdomain_coords[a, 1]:
                p.outside = True
                return p
    ref_lev_max = geom.refine_level_max  # :985
    ref_lev = min(refine_level, ref_lev_max)  # :995
    if ref_lev > 0:
        ref_lev_max = ref_lev  # :997-998
    p.ref_lev_max = int(ref_lev_max)
    bb = geom.block_bounds
    grid_bound_box = np.zeros_like(bb[0, ...])  # :1000-1002 (keeps the file dtype: f32 for plt)
    grid_bound_box[:, 0] = np.min(bb[..., 0], axis=0)
    grid_bound_box[:, 1] = np.max(bb[..., 1], axis=0)
    cellfac = 2 ** (ref_lev_max - 1)  # :1004
    grid_delta = (np.diff(grid_bound_box, axis=1).flatten() / (geom.nCellsVec * geom.nBlksVec * cellfac))[:, None]  # :1005-1007
    grid_half_delta = grid_delta * 0.5
    local_BCIDs = np.zeros((geom.nblocks, MESH_MDIM, 2), dtype=np.int32)  # :1010
    subdomain_BCIDs = np.zeros((MESH_MDIM, 2), dtype=np.int32)
    for lb in range(geom.nblocks):  # :1013-1015 (assignment into int32 truncates toward zero)
        local_BCIDs[lb, :, :] = (bb[lb] - grid_bound_box[:, 0, None] + grid_half_delta) / grid_delta
    if p.subdomain_flag:  # :1017-1022
        subdomain_BCIDs[:MESH_MDIM, :] = 0.5 + (subdomain_coords[:MESH_MDIM, :] - grid_bound_box[:MESH_MDIM, :1]) / grid_delta[:MESH_MDIM, :]
    max_scale = int(2 ** (ref_lev_max - 1))  # :1024
    fine_blks = max_scale * np.array(geom.nBlksVec, dtype=np.int32)  # :1026
    subd_cells = np.ones_like(fine_blks)
    if p.subdomain_flag:
        subd_cells[:ndim] = np.diff(subdomain_BCIDs[:ndim, :]).flatten()  # :1033-1034
    # (:1036-1154 is a per-rank pencil decomposition whose results are never used)
    local_BCIDs[:, ndim:MESH_MDIM, 1] = 0  # :1159 / :1175

    def intersects(lbc):  # _intersects_subdomain :1386-1393
        if not p.subdomain_flag:
            return True
        return all(subdomain_BCIDs[n, 0] <= lbc[n, 1] and lbc[n, 0] <= subdomain_BCIDs[n, 1] for n in range(MESH_MDIM))

    leaf_IDs = []
    for lb in range(geom.nblocks):  # get_blocklist("ALL") :1161 / :1176
        if ref_lev > -1:
            maybe = (geom.node_type[lb] == 1 and geom.refine_level[lb] < ref_lev) or geom.refine_level[lb] == ref_lev  # :1163-1165
        else:
            maybe = geom.node_type[lb] == 1  # :1177
        if maybe and intersects(local_BCIDs[lb, ...]):
            leaf_IDs.append(lb)
    if p.subdomain_flag:
        p.refdom_bound_box = grid_bound_box[:, :1] + subdomain_BCIDs * grid_delta  # :1185
        total_cells = np.copy(subd_cells)  # :1191
    else:
        p.refdom_bound_box = np.copy(grid_bound_box)  # :1188
        total_cells = np.ones_like(fine_blks)
        total_cells[:ndim] = fine_blks[:ndim] * geom.nCellsVec[:ndim]  # :1193-1194
    p.grid_delta = grid_delta
    p.local_BCIDs = local_BCIDs
    p.subdomain_BCIDs = subdomain_BCIDs
    p.leaf_IDs = leaf_IDs
    p.total_cells = total_cells
    return p


# ------------------------------------------------------------------------------------------------
# A9 — from_amr gather (_flash.py:1208-1321)
# ------------------------------------------------------------------------------------------------
def from_amr_gather_dict(geom: MeshGeom, plan: AmrPlan, field: np.ndarray) -> np.ndarray:
    """Literal restatement of the reference's per-cell dict mapping (:1262-1314).  Pure-Python loops:
    small cases only.  field = float64[blk,i,j,k]; returns float64[NX,NY,NZ]."""
    nxb, nyb, nzb = (int(v) for v in geom.nCellsVec)
    ndim = geom.ndim
    sd = plan.subdomain_BCIDs
    mapping = {}
    for leaf in plan.leaf_IDs:
        offx = plan.local_BCIDs[leaf, 0, 0]
        offy = plan.local_BCIDs[leaf, 1, 0] if ndim > 1 else 0
        offz = plan.local_BCIDs[leaf, 2, 0] if ndim > 2 else 0
        scale = int(2 ** (plan.ref_lev_max - geom.refine_level[leaf]))  # :1270-1271
        for i, j, k in itertools.product(range(nxb), range(nyb), range(nzb)):  # :1220-1222, :1273
            for ii, jj, kk in itertools.product(
                range(i * scale, (i + 1) * scale),
                range(j * scale if ndim > 1 else 0, (j + 1) * scale if ndim > 1 else 1),
                range(k * scale if ndim > 2 else 0, (k + 1) * scale if ndim > 2 else 1),
            ):
                I, J, K = offx + ii, offy + jj, offz + kk
                if plan.subdomain_flag:
                    inside = sd[0, 0] <= I < sd[0, 1] and sd[1, 0] <= J < sd[1, 1] and sd[2, 0] <= K < sd[2, 1]  # :1379-1384
                    if not inside:
                        continue
                    I, J, K = I - sd[0, 0], J - sd[1, 0], K - sd[2, 0]  # :1302
                mapping[(int(I), int(J), int(K))] = (leaf, i, j, k)  # :1305
    out = np.zeros(tuple(int(v) for v in plan.total_cells), dtype=np.float64)  # :1230, :1258
    for dest, src in mapping.items():  # :1313-1314
        out[dest] = field[src]
    return out


def from_amr_gather(geom: MeshGeom, plan: AmrPlan, field: np.ndarray) -> np.ndarray:
    """Vectorised equivalent of `from_amr_gather_dict` (np.repeat injection per leaf, in leaf-list order
    so later leaves overwrite earlier ones like the dict does).  Bit-identical; usable at 256^3."""
    sd = plan.subdomain_BCIDs
    tot = tuple(int(v) for v in plan.total_cells)
    out = np.zeros(tot, dtype=np.float64)
    for leaf in plan.leaf_IDs:
        scale = int(2 ** (plan.ref_lev_max - geom.refine_level[leaf]))
        blk = field[leaf]
        if scale > 1:
            blk = np.repeat(np.repeat(np.repeat(blk, scale, axis=0), scale, axis=1), scale, axis=2)
        elif scale < 1:
            raise ValueError("leaf finer than target level")
        off = [int(plan.local_BCIDs[leaf, a, 0]) for a in range(3)]
        lo = [0, 0, 0]
        hi = list(tot)
        if plan.subdomain_flag:
            off = [off[a] - int(sd[a, 0]) for a in range(3)]
        src = []
        dst = []
        empty = False
        for a in range(3):
            d0 = max(off[a], lo[a])
            d1 = min(off[a] + blk.shape[a], hi[a])
            if d1 <= d0:
                empty = True
                break
            dst.append(slice(d0, d1))
            src.append(slice(d0 - off[a], d1 - off[a]))
        if not empty:
            out[tuple(dst)] = blk[tuple(src)]
    return out


# ------------------------------------------------------------------------------------------------
# next-row: FLASH.slice_integral / slice_average (_flash.py:1427-1504)
# ------------------------------------------------------------------------------------------------
def slice_integral(geom: MeshGeom, field: np.ndarray, axis: int = 0):
    """Restatement of _flash.py:1451-1504 for field = float64[blk,i,j,k]; as in `reynolds_stress` the
    plane of constant index along `axis` is reduced (the reference always reduces array axes (1,2))."""
    ndim = geom.ndim
    lrefcells = 2 ** (geom.refine_level_max - 1)
    dims = [int(nb * bl * lrefcells) for nb, bl in zip(geom.nCellsVec[:ndim], geom.nBlksVec[:ndim])]
    min_delta = geom.min_delta(axis)
    db = geom.domain_bounds
    nrb = int(geom.nCellsVec[axis])
    span = np.linspace(db[axis, 0], db[axis, 1], dims[axis] + 1, dtype=np.float64)  # :1479
    blocklist = geom.leaf_blocks()
    alp = np.zeros(dims[axis], dtype=np.float64)
    cell_vols = np.array([geom.cell_volume_from_level(geom.refine_level[b]) for b in blocklist], dtype=np.float64)
    vol_fracs = cell_vols * (min_delta / geom.delta_from_level(axis, geom.refine_level[blocklist]))  # :1483-1486
    for lb, blk in enumerate(blocklist):  # :1488-1498
        lref_n = int(2 ** (geom.refine_level_max - 1) / 2 ** (geom.refine_level[blk] - 1))
        ilo = int(np.argmin(np.abs(span[:-1] - geom.block_bounds[blk, axis, 0])))
        mean = np.einsum("ijk->i", np.moveaxis(field[blk, ...], axis, 0)) * vol_fracs[lb]
        for i in range(nrb):
            alp[ilo + i * lref_n : ilo + (i + 1) * lref_n] += mean[i]
    return span, alp


def slice_average(geom: MeshGeom, field: np.ndarray, axis: int = 0):
    """_flash.py:1427-1449: slice_integral / (min_delta * cross-section)."""
    db = geom.domain_bounds
    others = [a for a in range(3) if a != axis]
    layer = (db[others[0], 1] - db[others[0], 0]) * (db[others[1], 1] - db[others[1], 0])
    span, alp = slice_integral(geom, field, axis)
    return span, alp / (geom.min_delta(axis) * layer)


# ------------------------------------------------------------------------------------------------
# §8f rank 4 — FlashUniform.fractal_dimension (fava/mesh/FLASH/FlashUniform.py:85-227)
# ------------------------------------------------------------------------------------------------
_NEIGHBOURS = ((1, 0, 0), (0, 1, 0), (0, -1, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1))  # order of :137-177


def fractal_marks_loop(d: np.ndarray, contour: float) -> np.ndarray:
    """Literal per-cell loop of FlashUniform.py:114-177 (small arrays only): int8[H,W,D] edge flags."""
    h, w, dp = d.shape
    e = np.zeros(d.shape, dtype=np.int8)
    e[d == contour] = 1  # :122
    for i, j, k in itertools.product(range(1, h - 1), range(1, w - 1), range(1, dp - 1)):  # :133-134
        val = d[i, j, k]
        if val < contour:
            hidx = contour - val
            for di, dj, dk in _NEIGHBOURS:
                nb = d[i + di, j + dj, k + dk]
                if nb > contour:
                    if int(hidx / (nb - val)) == 0:  # :142, :148, ... crossing nearer to the low cell
                        e[i, j, k] = 1
                    else:
                        e[i + di, j + dj, k + dk] = 1
    return e


def fractal_marks(d: np.ndarray, contour: float) -> np.ndarray:
    """Vectorised form of the same marking (order-independent: flags are only ever set)."""
    d = np.asarray(d, dtype=np.float64)
    h, w, dp = d.shape
    e = np.zeros(d.shape, dtype=np.int8)
    e[d == contour] = 1
    if min(h, w, dp) < 3:
        return e
    inner = (slice(1, h - 1), slice(1, w - 1), slice(1, dp - 1))
    val = d[inner]
    below = val < contour
    hidx = contour - val
    for di, dj, dk in _NEIGHBOURS:
        shifted = (slice(1 + di, h - 1 + di), slice(1 + dj, w - 1 + dj), slice(1 + dk, dp - 1 + dk))
        nb = d[shifted]
        cross = below & (nb > contour)
        with np.errstate(divide="ignore", invalid="ignore"):
            near_low = np.trunc(hidx / (nb - val)) == 0  # int(x) == 0
        e[inner][cross & near_low] = 1
        e[shifted][cross & ~near_low] = 1
    return e


def box_counts(e: np.ndarray) -> np.ndarray:
    """Filled boxes per level 0..flength-1, box edge 2^level (FlashUniform.py:179-208)."""
    h, w, dp = e.shape
    flength = int(np.log2(min(h, w, dp)) + 1)  # :184 with lowest_level = 0
    counts = np.zeros(flength, dtype=np.int64)
    for level in range(flength):
        b = 2**level
        if h % b or w % b or dp % b:
            raise IndexError(f"box edge {b} does not tile a {e.shape} grid (the reference indexes out of bounds here)")
        boxes = (e.reshape(h // b, b, w // b, b, dp // b, b) > 0).any(axis=(1, 3, 5))
        counts[level] = int(boxes.sum())
    return counts


def fractal_regression(counts: np.ndarray) -> dict:
    """FlashUniform.py:207-226 on the integer box counts: keys of one contour's result dict."""
    counts = np.asarray(counts)
    flength = counts.shape[0]
    result = np.zeros((flength, 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        for level in range(flength):
            result[level, 0] = flength - level - 1
            result[level, 1] = np.log2(counts[level])
        filled_boxes = 2 ** result[:, 1]
        cum_frac_dim = np.sum(np.log2(filled_boxes[:-1] / filled_boxes[1:]))
        avg_frac_dim = cum_frac_dim / (filled_boxes.size - 1.0)
        mean = np.mean(result, axis=0)
        std = np.std(result, axis=0)
        rval = np.sum((result[:, 0] - mean[0]) * (result[:, 1] - mean[1])) / (np.prod(std) * result.shape[0])
        slope = rval * std[1] / std[0]
        regress = np.array([slope, rval**2, mean[1] - slope * mean[0]])
    return {"average fractal dimension": avg_frac_dim, "slope": regress[0], "R2": regress[1], "curve": regress[2]}


def fractal_dimension(d: np.ndarray, field: str, contour: float) -> dict:
    """{field: {str(contour): {...}}} for a [i,j,k] float64 array (one float contour, as the reference accepts)."""
    return {field: {f"{contour}": fractal_regression(box_counts(fractal_marks(d, contour)))}}


# ------------------------------------------------------------------------------------------------
# §8f rank 4 — FlashUniform.structure_functions (FlashUniform.py:306-445)
# ------------------------------------------------------------------------------------------------
def structure_function_points(domain_bounds, sep: float, num_points: int):
    """One separation's random point pairs, consuming numpy's GLOBAL RandomState exactly like
    FlashUniform.py:361-395: 3*num_points uniforms (point 1), num_points (phi), num_points (theta); point 2 is
    wrapped periodically into the domain one period at a time."""
    db = np.asarray(domain_bounds, dtype=np.float64)
    ndim = db.shape[0]
    p1 = np.random.random((num_points, ndim)) * np.diff(db, axis=1).ravel() + db[:, 0].ravel()
    phi = 2.0 * np.pi * np.random.random(num_points)
    theta = np.arccos(2.0 * np.random.random(num_points) - 1.0)
    p2 = np.empty_like(p1)
    p2[:, 0] = p1[:, 0] + sep * np.sin(theta) * np.cos(phi)
    p2[:, 1] = p1[:, 1] + sep * np.sin(theta) * np.sin(phi)
    p2[:, 2] = p1[:, 2] + sep * np.cos(theta)
    for ax in range(3):
        lo, hi = db[ax]
        while np.any(p2[:, ax] > hi):
            p2[p2[:, ax] > hi, ax] += lo - hi
        while np.any(p2[:, ax] < lo):
            p2[p2[:, ax] < lo, ax] += hi - lo
    return p1, p2


def structure_functions(vel: dict, ncells_vec, domain_bounds, num_seps=100, num_points=10000, sep_bounds=(0.0, 1.0),
                        log_scale=True, anistropic=False) -> dict:
    """vel: {"velx","vely","velz"} float64 [i,j,k].  Orders 1..10, a fresh sample per order (FlashUniform.py:349)."""
    db = np.asarray(domain_bounds, dtype=np.float64)
    ncv = np.asarray(ncells_vec)
    names = ("velx", "vely", "velz")
    separations = np.geomspace(*sep_bounds, num_seps) if log_scale else np.linspace(*sep_bounds, num_seps)
    cell_size = np.diff(db, axis=1).flatten() / ncv
    out = {"transverse": {}, "longitudinal": {}}
    for order in range(1, 11):
        pt = np.zeros((num_seps, num_points, 3, 2))
        dv = np.zeros((num_seps, num_points, 3))
        for i in range(num_seps):
            p1, p2 = structure_function_points(db, separations[i], num_points)
            i1 = [np.floor((p1[:, j] - db[j, 0]) / cell_size[j]).astype(int) for j in range(3)]
            i2 = [np.floor((p2[:, j] - db[j, 0]) / cell_size[j]).astype(int) for j in range(3)]
            for j, name in enumerate(names):
                dv[i, :, j] = vel[name][i2[0], i2[1], i2[2]] - vel[name][i1[0], i1[1], i1[2]]
            pt[i, ..., 0] = p1
            pt[i, ..., 1] = p2
        sep_vec = pt[..., 1] - pt[..., 0]
        rhat = np.empty_like(sep_vec)
        if anistropic:
            rhat[..., 0] = 1.0
            rhat[..., 1:] = 0.0
        else:
            for j in range(3):
                rhat[..., j] = sep_vec[..., j] / np.sqrt(np.sum(sep_vec**2, axis=2))
        long_comp = np.abs(np.sum(dv * rhat, axis=2))
        out["longitudinal"][f"{order}"] = np.sum(long_comp**order, axis=1) / float(num_points)
        long_dvel = long_comp[..., None] * rhat
        trans_comp = np.sqrt(np.sum((dv - long_dvel) ** 2, axis=2))
        out["transverse"][f"{order}"] = np.sum(trans_comp**order, axis=1) / float(num_points)
        out["separations"] = separations
    return out
