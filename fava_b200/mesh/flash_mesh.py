"""FLASH block-structured mesh on the GPU — the drop-in for the reference's `fava.mesh.FLASH.FLASH`
(fava/mesh/FLASH/_flash.py) on the grid-statistics hot path.

Same surface: `FLASH(filename).load()`, `.load_data(names)`, `.data(name)`, `.reynolds_stress(raxis=0)`,
`.slice_integral/.slice_average(field, axis)`, `.from_amr(subdomain_coords, refine_level, fields, filename)`,
`.save(filename, names)` and the metadata attributes (`ndim, nxb, nyb, nzb, nblocks, xmin..zmax,
domain_bounds, nCellsVec, nBlksVec, fields, time, block_bounds, refine_level, node_type, ...`).

What differs underneath:
  * field data lives in HBM in the FILE layout and FILE dtype ([block][z][y][x], f32 for plt files): the
    dataset's byte range (h5lite index) is `pread` into a pinned ring and copied with async H2D copies
    (fava_stage_h2d).  The reference's float64 [blk, i, j, k] array (_flash.py:331-335) is only
    materialised on the host if a caller asks for `.data(name)`;
  * every reduction / gather runs in libfava_b200's sm_100a kernels through the C ABI; host code only
    builds the O(nblocks) tables with the reference's own arithmetic.  There is no CPU fallback;
  * ranks (one process per GPU, torch.distributed) own contiguous block ranges exactly like
    `_mpi_assign_blocks` (_flash.py:166-208) — or z-slabs of a single-block dataset — and exchange one
    packed [14][N] fp64 all-reduce instead of ten Allreduce calls (_flash.py:1581, :1608).

Deliberate deviations from the reference's *behaviour* (SURVEY §0): `reynolds_stress` accepts `axis=` as an
alias of `raxis=` (README.rst:31) and for axis != 0 returns the profile ALONG THAT AXIS (the reference still
reduces over array axes (1,2), _flash.py:1570, and so returns the x-profile labelled with y/z coordinates);
`from_amr` also accepts lists and the long field names of FIELD_MAPPING.
"""

from __future__ import annotations

import logging
from functools import cached_property
from pathlib import Path

import numpy as np
import torch

from fava_b200 import device, dist, h5lite, stats
from fava_b200.geometry import AXIS, GEOMETRY
from fava_b200.mesh.base import Structured
from fava_b200.model import Model
from fava_b200.util import HID_T, timer

logger = logging.getLogger(__name__)

# long name -> 4-character FLASH dataset name (reference fava/mesh/FLASH/_util.py:1-13)
FIELD_MAPPING = {
    "velocity-x": "velx",
    "velocity-y": "vely",
    "velocity-z": "velz",
    "density": "dens",
    "pressure": "pres",
    "temperature": "temp",
    "energy": "ener",
    "flame progress": "flam",
    "ignition time": "igtm",
    "velocity-divergence": "divv",
    "vorticity": "vort",
}
NGUARD = 4
MESH_MDIM = 3
LEAF = 1  # node type of a leaf block (BLOCK_TYPE.LEAF, _flash.py:28-29)

_STRESS_KEYS = ("Rxx", "Rxy", "Rxz", "Ryy", "Ryz", "Rzz")
_VEL = ("velx", "vely", "velz")


def _strip(a) -> np.ndarray:
    return np.char.strip(np.asarray(a).astype(str))


@Model.register_mesh()
class FLASH(Structured):
    """FLASH AMR / multi-block mesh backed by device-resident block data."""

    def __init__(self, filename=None, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._filename: Path | None = None
        self._chk_file = False
        self._dev: dict[str, torch.Tensor] = {}  # field -> device tensor, FILE layout
        self._host_cache: dict[str, np.ndarray] = {}
        self._extent: dict[str, tuple] = {}  # field -> (offset, nbytes, dtype, shape) in the file
        self._part = None  # ("blocks", b0, b1) | ("slab", z0, z1) of what this rank holds
        self.fields = np.array([], dtype=str)
        self.scalars = {"real": {}, "integer": {}, "logical": {}, "string": {}}
        self.runtime_parameters = {"real": {}, "integer": {}, "logical": {}, "string": {}}
        self.filename = filename

    # ---- identity ---------------------------------------------------------------------------------
    @classmethod
    def is_this_your_mesh(cls, filename, *args, **kwargs) -> bool:
        return any(tag in str(filename) for tag in ("hdf5_chk_", "hdf5_plt_cnt_"))

    @property
    def filename(self) -> Path | None:
        return self._filename

    @filename.setter
    def filename(self, filename) -> None:
        if not isinstance(filename, (str, Path)):
            if filename is not None:
                logger.error("Filename must be passed in as a str or Path; not %s", type(filename))
            return
        fn = Path(filename)
        if fn == self._filename:
            return
        self._filename = fn
        if "chk" in fn.stem:  # checkpoint files carry f64 reals (_flash.py:80-81)
            self._chk_file = True

    # ---- metadata ---------------------------------------------------------------------------------
    _META_ALL = ("bflags", "coordinates", "block size", "bounding box", "processor number", "node type",
                 "refine level", "gid", "which child")

    def load(self) -> None:
        """Everything but the field arrays (reference FLASH.load, _flash.py:106-163)."""
        if self._filename is None or not self._filename.is_file():
            logger.error("File does not exist: %s", self._filename)
            return
        try:
            self._reset_data()
            with h5lite.File(self._filename, "r") as f:
                self._read_parameters(f)
                self._set_integers()
                self._set_reals()
                if self.nblocks < 1:
                    raise ValueError("[FLASH reader] the file reports no blocks")
                self.fields = np.squeeze(f["unknown names"][()]).astype(str).reshape(-1)
                self._read_block_metadata(f, self._META_ALL)
                self._index_fields(f)
        except Exception as exc:
            logger.exception("Error reading FLASH FILE %s", self._filename)
            raise RuntimeError(f"Error reading FLASH FILE {self._filename}") from exc
        self._loaded = True

    def _reset_data(self) -> None:
        self._dev.clear()
        self._host_cache.clear()
        self._extent.clear()
        self._part = None
        for name in ("geometry", "refine_level_max", "domain_volume"):
            self.__dict__.pop(name, None)

    def _read_parameters(self, f) -> None:
        """The eight compound datasets {name, value} (reference _read_scalars / _read_runtime_parameters,
        _flash.py:211-252): all four kinds must exist."""
        for group, suffix in ((self.scalars, "scalars"), (self.runtime_parameters, "runtime parameters")):
            for key in ("real", "integer", "logical", "string"):
                rows = f[f"{key} {suffix}"][()]
                names = _strip(rows["name"]) if rows.size else np.array([], dtype=str)
                values = _strip(rows["value"]) if key == "string" and rows.size else rows["value"]
                group[key] = dict(zip(names, values))

    def _set_integers(self) -> None:  # _flash.py:379-393
        si, ri = self.scalars["integer"], self.runtime_parameters["integer"]
        self.ndim = np.int64(si.get("dimensionality"))
        self.nxb, self.nyb, self.nzb = (np.int64(si.get(k)) for k in ("nxb", "nyb", "nzb"))
        self.iprocs, self.jprocs, self.kprocs = (np.int64(si.get(k, 1)) for k in ("iprocs", "jprocs", "kprocs"))
        self.nblockx, self.nblocky, self.nblockz = (np.int64(ri.get(k)) for k in ("nblockx", "nblocky", "nblockz"))
        self.nblocks = np.int64(si.get("total blocks", si.get("globalnumblocks")))

    def _set_reals(self) -> None:  # _flash.py:370-377
        self.time = np.float64(self.scalars["real"].get("time"))
        rr = self.runtime_parameters["real"]
        self.xmin, self.xmax = np.float64(rr.get("xmin", 0)), np.float64(rr.get("xmax", 1))
        self.ymin, self.ymax = np.float64(rr.get("ymin", 0)), np.float64(rr.get("ymax", 1))
        self.zmin, self.zmax = np.float64(rr.get("zmin", 0)), np.float64(rr.get("zmax", 1))

    _ATTR = {"bflags": "bflags", "coordinates": "coordinates", "block size": "block_size",
             "bounding box": "block_bounds", "processor number": "processors", "node type": "node_type",
             "refine level": "refine_level", "gid": "gid", "which child": "which_child"}
    _AS_INT64 = ("processor number", "node type", "refine level", "gid", "which child")

    def _read_block_metadata(self, f, keys) -> None:
        """Block tables; reals keep the file dtype (f32 for plt), ints widen to int64 (_flash.py:263-304)."""
        for key in keys:
            arr = f[key][()]
            if key in self._AS_INT64:
                arr = arr.astype(np.int64)
            setattr(self, self._ATTR[key], arr)

    def _index_fields(self, f) -> None:
        for name in self.fields:
            key = f"{name:4s}"
            if key in f:
                d = f[key]
                off, nbytes = d.extent()
                self._extent[str(name)] = (off, nbytes, d.dtype, d.shape)

    # ---- derived geometry (host scalars; reference _flash.py:396-411, :583-602, :914-953) ----------
    @property
    def domain_bounds(self) -> np.ndarray:
        return np.array([[self.xmin, self.xmax], [self.ymin, self.ymax], [self.zmin, self.zmax]], dtype=np.float64)

    @property
    def ncells(self) -> int:
        return self.nxb * self.nyb * self.nzb

    @property
    def nCellsVec(self) -> np.ndarray:
        return np.array([self.nxb, self.nyb, self.nzb], dtype=np.int32)

    @property
    def nBlksVec(self) -> np.ndarray:
        return np.array([self.nblockx, self.nblocky, self.nblockz], dtype=np.int32)

    @cached_property
    def geometry(self) -> GEOMETRY:
        return GEOMETRY(str(self.scalars["string"].get("geometry", "")).lower())

    @cached_property
    def refine_level_max(self) -> int:
        return self.refine_level.max()

    @cached_property
    def domain_volume(self) -> float:
        if self.geometry is not GEOMETRY.CARTESIAN:
            raise NotImplementedError(f"Domain volume not implemented for {self.geometry}")
        return np.prod(np.diff(self.domain_bounds))

    def get_minimum_deltas(self, axis: int):
        d = self.domain_bounds
        return (d[axis, 1] - d[axis, 0]) / (self.nCellsVec[axis] * self.nBlksVec[axis] * 2 ** (self.refine_level_max - 1))

    def get_delta_from_refine_level(self, axis: int, refine_level):
        d = self.domain_bounds
        return (d[axis, 1] - d[axis, 0]) / (self.nCellsVec[axis] * self.nBlksVec[axis] * 2 ** (refine_level - 1))

    def get_cell_volume_from_refinement(self, refine_level=1):
        cells = self.nxb * self.nblockx * 2 ** (refine_level - 1)
        if self.ndim > 1:
            cells = cells * (self.nyb * self.nblocky * 2 ** (refine_level - 1))
        if self.ndim > 2:
            cells = cells * (self.nzb * self.nblockz * 2 ** (refine_level - 1))
        return self.domain_volume / np.asarray(cells, dtype=np.float64)

    # ---- block ownership (reference _mpi_assign_blocks & co, _flash.py:166-208) ---------------------
    @property
    def blk_beg(self) -> int:
        return dist.parallel_range(int(self.nblocks))[0]

    @property
    def blk_end(self) -> int:
        return dist.parallel_range(int(self.nblocks))[1]

    @property
    def nblocks_local(self) -> int:
        return self.blk_end - self.blk_beg

    def get_blocklist(self, block_type: str = "LEAF") -> np.ndarray:
        lb, ub = self.blk_beg, self.blk_end
        name = block_type if isinstance(block_type, str) else getattr(block_type, "name", str(block_type))
        if name == "LEAF":
            return (lb + np.flatnonzero(self.node_type[lb:ub] == LEAF)).astype(np.int64)
        if name == "ALL":
            return np.arange(lb, ub, dtype=np.int64)
        logger.error("Do not recognize BLOCK TYPE %s", name)
        raise ValueError(name)

    def get_cell_volumes(self, block_type="LEAF") -> np.ndarray:
        return self.get_cell_volume_from_refinement(self.refine_level[self.get_blocklist(block_type)])

    # ---- field data ---------------------------------------------------------------------------------
    def _resolve(self, name: str) -> str | None:
        if name in self._dev or name in self._extent or name in self.fields:
            return str(name)
        return FIELD_MAPPING.get(name)

    def _partition_of(self, shape) -> tuple:
        """Which part of a dataset this rank stages: its contiguous block range of a multi-block
        dataset, or a z-slab of a single-block / 3-D one.  Both are contiguous byte ranges."""
        if len(shape) == 4 and shape[0] > 1:
            b0, b1 = dist.parallel_range(int(shape[0]))
            return ("blocks", b0, b1)
        nz = int(shape[-3])
        z0, z1 = dist.parallel_range(nz)
        return ("slab", z0, z1)

    def load_data(self, names=None) -> None:
        """Stage fields into HBM (reference load_data / _read_variable_data, _flash.py:84-88, :306-341)."""
        for field in (names if names is not None else self.fields):
            key = self._resolve(field)
            if key is None:
                raise KeyError(f"{field} field not found in dataset {self._filename}")
            self._stage(key)

    def _stage(self, key: str) -> torch.Tensor:
        if key in self._dev:
            return self._dev[key]
        if key not in self._extent:
            raise RuntimeError(f"Error occurred in reading FLASH VARIABLE DATA: {key!r} field not found in dataset "
                               f"{self._filename}")
        off, nbytes, dtype, shape = self._extent[key]
        if dtype not in (np.dtype("<f4"), np.dtype("<f8")):
            raise RuntimeError(f"field {key!r} has dtype {dtype}; FLASH fields are little-endian f32 or f64")
        part = self._partition_of(shape)
        if self._part is None:
            self._part = part
        elif self._part != part:
            raise RuntimeError("fields of one file must share a shape")
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if dev is None:
            raise RuntimeError("fava_b200 needs a CUDA device (B200): field data is staged into HBM and there is "
                               "no CPU fallback")
        if part[0] == "blocks":
            per = int(np.prod(shape[1:])) * dtype.itemsize
            lshape = (part[2] - part[1],) + tuple(shape[1:])
        else:
            per = int(np.prod(shape[-2:])) * dtype.itemsize
            lshape = (part[2] - part[1],) + tuple(shape[-2:])
        out = torch.empty(lshape, dtype=tdt, device=dev)
        device.stage_file(self._filename, off + part[1] * per, (part[2] - part[1]) * per, out)
        self._dev[key] = out
        return out

    def _stage_block_range(self, key: str, b0: int, b1: int) -> torch.Tensor:
        """Blocks [b0, b1) of a 4-D field dataset straight from the file (contiguous byte range)."""
        if key not in self._extent:
            raise RuntimeError(f"field {key!r} has no on-disk block dataset to stage from")
        off, nbytes, dtype, shape = self._extent[key]
        if len(shape) != 4:
            raise RuntimeError("from_amr on several ranks needs a block dataset [nblocks][nzb][nyb][nxb]")
        per = int(np.prod(shape[1:])) * dtype.itemsize
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        out = torch.empty((b1 - b0,) + tuple(shape[1:]), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
        device.stage_file(self._filename, off + b0 * per, (b1 - b0) * per, out)
        return out

    def _stage_plane_range(self, key: str, z0: int, z1: int) -> torch.Tensor:
        """Planes [z0, z1) of a single-block / 3-D field dataset straight from the file (contiguous byte range)."""
        if key not in self._extent:
            raise RuntimeError(f"field {key!r} has no on-disk dataset to stage from")
        off, nbytes, dtype, shape = self._extent[key]
        if len(shape) == 4 and shape[0] != 1:
            raise RuntimeError("plane ranges exist for single-block datasets only")
        ny, nx = int(shape[-2]), int(shape[-1])
        per = ny * nx * dtype.itemsize
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        out = torch.empty((z1 - z0, ny, nx), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
        device.stage_file(self._filename, off + z0 * per, (z1 - z0) * per, out)
        return out

    def device_data(self, name: str) -> torch.Tensor:
        """Device tensor of this rank's part of a field, FILE layout ([block][z][y][x] or [z][y][x])."""
        key = self._resolve(name)
        if key is None:
            raise KeyError(name)
        return self._stage(key)

    def data(self, name: str):
        """The reference's view of a field: float64, axes -1/-3 swapped ([blk, i, j, k] / [i, j, k],
        _flash.py:90-104, :331-335) — a HOST copy made on demand; kernels never use it."""
        key = self._resolve(name)
        if key is None:
            logger.warning("Cannot find %s in dataset", name)
            return None
        if key not in self._host_cache:
            t = self._stage(key)
            host = t.to(torch.float64).cpu().numpy()
            if t.dim() == 3 and len(self._extent.get(key, (0, 0, 0, (0,) * 3))[3]) == 4:
                host = host[None, ...]
            self._host_cache[key] = np.ascontiguousarray(np.swapaxes(host, -1, -3))
        return self._host_cache[key]

    @property
    def _data(self) -> dict:
        """Mapping view used like the reference's `_data` dict (keys = loaded fields)."""
        return {k: self.data(k) for k in self._dev}

    # ---- plane statistics ------------------------------------------------------------------------
    def _axis_setup(self, axis: int):
        ax = AXIS(axis)  # ValueError for anything but 0,1,2 (reference :1513)
        lrefcells = 2 ** (self.refine_level_max - 1)
        dims = [int(nb * bl * lrefcells) for nb, bl in zip(self.nCellsVec[: self.ndim], self.nBlksVec[: self.ndim])]
        min_delta = self.get_minimum_deltas(ax.value)
        b = self.domain_bounds
        others = [a for a in range(3) if a != ax.value]
        layer_volume = (b[others[0], 1] - b[others[0], 0]) * (b[others[1], 1] - b[others[1], 0]) * min_delta
        nrb = int(self.nCellsVec[ax.value])
        n = dims[ax.value]
        radius = np.linspace(b[ax.value, 0], b[ax.value, 1], n + 1)
        return ax.value, n, nrb, float(min_delta), float(layer_volume), radius

    def _leaf_bins(self, axis: int, radius: np.ndarray, blocklist: np.ndarray):
        """(ilo, lref_n, vol_frac) per leaf with the reference's arithmetic (_flash.py:1559-1567):
        ilo = argmin |radius[:-1] - bbox_lo| (first minimum), lref_n = 2^(lmax-level),
        vol_frac = V_cell(level) * dmin / d(level)."""
        lev = self.refine_level[blocklist]
        scale = (2 ** (self.refine_level_max - 1) / 2 ** (lev - 1)).astype(np.int64)
        lo = self.block_bounds[blocklist, axis, 0]
        edges = radius[:-1]
        j = np.clip(np.searchsorted(edges, lo), 1, max(edges.size - 1, 1))
        if edges.size > 1:
            left_closer = np.abs(edges[j - 1] - lo) <= np.abs(edges[j] - lo)
            ilo = np.where(left_closer, j - 1, j)
        else:
            ilo = np.zeros(lo.shape, dtype=np.int64)
        vol = self.get_cell_volume_from_refinement(lev) * (
            self.get_minimum_deltas(axis) / self.get_delta_from_refine_level(axis, lev))
        return ilo.astype(np.int64), scale, np.asarray(vol, dtype=np.float64)

    def _plane_statistics(self, axis: int, favre: bool):
        if int(self.ndim) != 3:
            raise NotImplementedError("fava_b200 computes plane statistics for 3-D datasets (rho, velx, vely, velz)")
        axis, n, nrb, min_delta, layer_volume, radius = self._axis_setup(axis)
        t = [self.device_data(k) for k in ("dens",) + _VEL]
        if self._part[0] == "slab":
            fields = [x.reshape(x.shape[-3:]) for x in t]
            out = stats.slab_profiles(*fields, axis, float(self.get_cell_volume_from_refinement(1)), layer_volume,
                                      favre=favre, gather=True)
        else:
            blocklist = self.get_blocklist("LEAF")
            ilo, scale, vol = self._leaf_bins(axis, radius, blocklist)
            table = device.leaf_table(blocklist - self.blk_beg, ilo, scale, vol)
            out = stats.block_profiles(*t, axis, table, n, layer_volume, favre=favre)
        return radius, {k: v.cpu().numpy() for k, v in out.items()}

    @timer
    def reynolds_stress(self, raxis: int = 0, axis: int | None = None):
        """Plane profiles of <rho u'_i u'_j> about the volume-averaged plane means (reference
        _flash.py:1506-1611).  Returns (radius[N+1], {Rxx,Rxy,Rxz,Ryy,Ryz,Rzz}[N], {dens,velx,vely,velz}[N])."""
        radius, out = self._plane_statistics(raxis if axis is None else axis, favre=False)
        stress = {k: out["reynolds"][i] for i, k in enumerate(_STRESS_KEYS)}
        means = {k: out["means"][i] for i, k in enumerate(("dens",) + _VEL)}
        return radius, stress, means

    @timer
    def favre_stress(self, axis: int = 0):
        """Favre statistics from the same single pass (extension; not in the reference):
        (radius, {Rij: <rho u''_i u''_j>}, {"dens": <rho>, "velx": u~_x, ...})."""
        radius, out = self._plane_statistics(axis, favre=True)
        stress = {k: out["favre"][i] for i, k in enumerate(_STRESS_KEYS)}
        means = {"dens": out["means"][0]}
        means.update({k: out["favre_means"][i] for i, k in enumerate(_VEL)})
        return radius, stress, means

    def slice_integral(self, field: str, axis: int = 0):
        """Volume-weighted plane integral of one field (reference _flash.py:1451-1504)."""
        axis, n, nrb, min_delta, layer_volume, span = self._axis_setup(axis)
        t = self.device_data(field)
        if self._part[0] == "slab":
            out = stats.slab_plane_sum(t.reshape(t.shape[-3:]), axis, float(self.get_cell_volume_from_refinement(1)))
        else:
            blocklist = self.get_blocklist("LEAF")
            ilo, scale, vol = self._leaf_bins(axis, span, blocklist)
            out = stats.block_plane_sum(t, axis, device.leaf_table(blocklist - self.blk_beg, ilo, scale, vol), n)
        return span, out.cpu().numpy()

    def slice_average(self, field: str, axis: int = 0):
        """slice_integral / layer volume (reference _flash.py:1427-1449)."""
        ax, n, nrb, min_delta, layer_volume, _ = self._axis_setup(axis)
        span, alp = self.slice_integral(field, axis=ax)
        return span, alp / layer_volume

    @timer
    def flame_window(self, radius: np.ndarray, stress: dict, mask=None) -> float:
        """Centre of the flame brush: Levenberg-Marquardt fit of a super-Gaussian amp*exp(-2((x-x0)/sigma)^10) to
        Ryy + Rzz over the masked cells (host-side SciPy; reference _flash.py:1613-1659, same scalings)."""
        import scipy.optimize

        def super_gaussian(x, amp, x0, sigma):
            return amp * np.exp(-2 * ((x - x0) / sigma) ** 10)

        ma = mask if mask is not None else np.where(radius < np.inf)[0]
        xfact = 1.0e5
        rspan = radius[ma] / xfact
        rmin = np.min(rspan)
        rsyyzz = stress["Ryy"][ma] + stress["Rzz"][ma]
        rsyyzz = rsyyzz / 10.0 ** np.max(np.floor(np.log10(rsyyzz)))
        # like the reference the abscissa is shifted by rmin while the initial guess and the result are not
        opt, _ = scipy.optimize.curve_fit(super_gaussian, rspan - rmin, rsyyzz, method="lm",
                                          p0=(np.max(rsyyzz), rspan[np.argmax(rsyyzz)], np.std(rsyyzz)))
        return opt[1] * xfact

    # ---- AMR -> uniform ----------------------------------------------------------------------------
    def from_amr(self, subdomain_coords=None, refine_level: int = -1, fields=None, filename=None) -> None:
        """Prolong the AMR mesh onto a uniform (sub)domain at `refine_level` (-1 = finest), replace the
        selected fields by the uniform arrays, turn the mesh into a one-block mesh and write the uniform
        file (reference FLASH.from_amr, _flash.py:955-1377)."""
        from fava_b200.mesh.amr_plan import build_plan

        sd = np.asarray(subdomain_coords, dtype=np.float64) if subdomain_coords is not None else None
        if sd is None:
            raise TypeError("'NoneType' object is not iterable")  # what `any(... for sdc in None)` raises
        plan = build_plan(self, sd, int(refine_level))
        if plan is None:  # sub-domain pokes outside the domain: the reference returns silently (:967-977)
            return
        names = list(fields) if fields is not None else [str(f) for f in self.fields]
        keys = []
        for nm in names:
            key = self._resolve(nm)
            if key is None:
                raise KeyError(nm)
            keys.append(key)
        nx, ny, nz = (int(v) for v in plan.total_cells)
        # every rank fills its own z-slab of the uniform array from the leaves that touch it; with more
        # than one rank those source blocks are staged as one contiguous block-id range of the file
        z0, z1 = dist.parallel_range(nz)
        ids, off, scales = plan.slab_leaves(z0, z1, int(self.nzb))
        off = off - np.array([0, 0, z0], dtype=np.int64)[None, :]
        new_dev = {}
        for key in keys:
            if dist.world_size() == 1:
                blocks = self.device_data(key)
                if blocks.dim() == 3:
                    blocks = blocks[None, ...]
                base = 0
            else:
                base, end = (int(ids.min()), int(ids.max()) + 1) if ids.size else (0, 1)
                blocks = self._stage_block_range(key, base, end)
            if z1 == z0:  # more ranks than output planes
                new_dev[key] = torch.empty((0, ny, nx), dtype=torch.float64, device=blocks.device)
                continue
            table = device.prolong_table(ids - base, off, scales)
            new_dev[key] = device.prolong(blocks, table, (z1 - z0, ny, nx))
        # ---- the mesh becomes a single uniform block (:1340-1361) ----
        gd = plan.grid_delta
        self._dev = new_dev
        self._host_cache.clear()
        self._part = ("slab",) + dist.parallel_range(nz)
        self._extent = {}
        self.gid = -1 * np.ones(int(2 * self.ndim + 1 + 2**self.ndim), dtype=np.int32)
        self.refine_level = np.ones(1, dtype=np.int32)
        self.node_type = np.ones_like(self.refine_level)
        self.bflags = -1 * np.ones_like(self.refine_level)
        self.which_child = np.copy(self.which_child)
        self.nblockx = self.nblocky = self.nblockz = 1
        self.nblocks = 1
        self.nxb, self.nyb, self.nzb = plan.total_cells[0], plan.total_cells[1], plan.total_cells[2]
        self.block_size = (plan.total_cells * gd)[None, ...]
        self.block_bounds = plan.refdom_bound_box[None, ...]
        self.coordinates = (0.5 * np.sum(plan.refdom_bound_box, axis=1))[None, ...]
        self.xmin, self.xmax = plan.refdom_bound_box[0]
        self.ymin, self.ymax = plan.refdom_bound_box[1]
        self.zmin, self.zmax = plan.refdom_bound_box[2]
        for name in ("refine_level_max", "domain_volume"):
            self.__dict__.pop(name, None)
        self._sync_parameter_tables()
        if filename is None:
            stem = self._filename.stem.replace("plt_cnt", "uniform").replace("chk", "uniform")
            uni = self._filename.with_stem(stem)
        else:
            uni = Path(filename)
        self.save(filename=uni, names=keys)

    def _sync_parameter_tables(self) -> None:
        """The reference's property setters mirror nxb/nblockx/xmin... into the scalar / runtime-parameter
        dicts when the key exists there (_flash.py:413-566), so that `save` writes the new values."""
        for name in ("nxb", "nyb", "nzb", "nblockx", "nblocky", "nblockz"):
            for table in (self.scalars["integer"], self.runtime_parameters["integer"]):
                if name in table:
                    table[name] = getattr(self, name)
        for table in (self.scalars["integer"], self.runtime_parameters["integer"]):
            if "globalnumblocks" in table:  # the only block-count key the reference mirrors (:489-494)
                table["globalnumblocks"] = self.nblocks
        for name in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"):
            for table in (self.scalars["real"], self.runtime_parameters["real"]):
                if name in table:
                    table[name] = getattr(self, name)

    # ---- writer --------------------------------------------------------------------------------------
    def save(self, filename=None, names=None) -> None:
        """Write a FLASH-format file (reference FLASH.save, _flash.py:619-799): parameters as S256
        compounds, block tables, and the fields in FILE order — f32 unless this is a checkpoint."""
        target = Path(filename) if filename is not None else self._filename
        names_ = list(names) if names is not None else list(self._dev)
        # every rank contributes its part of each field to the root (z-slabs / block ranges in rank order)
        payload = {}
        for var in names_:
            if var not in self._dev:
                continue
            full = dist.gather_cat_to_root(self._dev[var])
            if dist.is_root():
                payload[var] = full.cpu().numpy()
        if not dist.is_root():
            return
        real = HID_T.F64 if self._chk_file else HID_T.F32
        try:
            with h5lite.File(target, "w") as f:
                self._write_parameters(f)
                f.create_dataset(name="coordinates", shape=self.coordinates.shape, dtype=real, data=self.coordinates)
                f.create_dataset(name="block size", shape=self.block_size.shape, dtype=real, data=self.block_size)
                f.create_dataset(name="bounding box", shape=self.block_bounds.shape, dtype=real, data=self.block_bounds)
                for key, attr in (("node type", "node_type"), ("refine level", "refine_level"), ("gid", "gid"),
                                  ("which child", "which_child")):
                    arr = getattr(self, attr)
                    f.create_dataset(name=key, shape=arr.shape, dtype=HID_T.I32, data=arr.astype(np.int32))
                f.create_dataset(name="bflags", shape=self.bflags.shape, dtype=HID_T.I32, data=self.bflags)
                fld = np.bytes_(names_)
                f.create_dataset(name="unknown names", shape=fld.shape, dtype=HID_T.UNKNOWN_NAMES, data=fld)
                for var, arr in payload.items():
                    f.create_dataset(name=var, shape=arr.shape, dtype=real, data=arr)
        except Exception as exc:
            logger.exception("Error saving FLASH FILE %s", target)
            raise RuntimeError(f"Error saving FLASH FILE {target}") from exc

    def _write_parameters(self, f) -> None:
        """Scalars and runtime parameters (_flash.py:657-719), including its quirk: BOTH string datasets
        receive the runtime-parameter rows."""
        kinds = {"real": HID_T.F64_PARAMETER, "integer": HID_T.I32_PARAMETER, "logical": HID_T.BOOL_PARAMETER}
        for key in self.scalars:
            if key == "string":
                rows = [(f"{k:256s}", f"{v:256s}") for k, v in self.runtime_parameters[key].items()]
                for suffix in ("runtime parameters", "scalars"):
                    f.create_dataset(name=f"{key} {suffix}", shape=len(rows), dtype=HID_T.STR_PARAMETER, data=rows)
                continue
            if key not in kinds:
                logger.warning("Do not recognize parameter set %s", key)
                continue
            for suffix, table in (("runtime parameters", self.runtime_parameters), ("scalars", self.scalars)):
                rows = [(f"{k:256s}", v) for k, v in table[key].items()]
                f.create_dataset(name=f"{key} {suffix}", shape=len(rows), dtype=kinds[key], data=rows)
