"""Device-side plumbing above the C ABI: contexts, tensor hand-off, and thin functional wrappers.

PyTorch is used only for device memory, streams and (elsewhere) torch.distributed; every number is
produced by libfava_b200's own kernels through ctypes.  All wrappers take CUDA tensors in the FLASH
file layout ([z][y][x], x fastest) and enqueue on torch's current stream.
"""

from __future__ import annotations

import ctypes as C
import itertools

import numpy as np
import torch

from fava_b200 import _lib
from fava_b200._lib import FAVA_F32, FAVA_F64, FAVA_NMOM

STRESS_KEYS = ("Rxx", "Rxy", "Rxz", "Ryy", "Ryz", "Rzz")
MEAN_KEYS = ("dens", "velx", "vely", "velz")


class Context:
    """One fava_ctx per CUDA device (owns workspaces, cuFFT plans, the pinned staging ring)."""

    def __init__(self, device: int):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("fava_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                               "there is no CPU fallback")
        h = C.c_void_p()
        _lib.check(lib.fava_init(int(device), C.byref(h)), "fava_init")
        self.lib = lib
        self.handle = h
        self.device = int(device)

    def close(self) -> None:
        if self.handle:
            self.lib.fava_shutdown(self.handle)
            self.handle = C.c_void_p()


_contexts: dict[int, Context] = {}


def get_context(device: int | torch.device | None = None) -> Context:
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    elif isinstance(device, torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
    else:
        idx = int(device)
    ctx = _contexts.get(idx)
    if ctx is None:
        ctx = _contexts[idx] = Context(idx)
    return ctx


def shutdown() -> None:
    for ctx in _contexts.values():
        ctx.close()
    _contexts.clear()


def launch_count() -> int:
    return int(_lib.load().fava_launch_count())


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return FAVA_F32
    if t.dtype == torch.float64:
        return FAVA_F64
    raise TypeError(f"field dtype must be float32 or float64, not {t.dtype}")


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _check_fields(rho, ux, uy, uz) -> tuple[int, int, int]:
    if rho.dim() != 3:
        raise ValueError(f"expected a [z][y][x] array, got shape {tuple(rho.shape)}")
    for t in (rho, ux, uy, uz):
        if not t.is_cuda:
            raise ValueError("fields must be CUDA tensors (no CPU fallback)")
        if t.shape != rho.shape or t.dtype != rho.dtype or t.device != rho.device:
            raise ValueError("rho, ux, uy, uz must share shape, dtype and device")
        if not t.is_contiguous():
            raise ValueError("fields must be contiguous in [z][y][x] order")
    nz, ny, nx = (int(s) for s in rho.shape)
    return nz, ny, nx


def plane_pivots(ux, uy, uz, axis: int) -> torch.Tensor:
    """[3][nbins] pivots: velocity at the first cell of each plane normal to `axis`."""
    nz, ny, nx = _check_fields(ux, ux, uy, uz)
    nbins = (nx, ny, nz)[axis]
    ctx = get_context(ux.device)
    piv = torch.empty((3, nbins), dtype=torch.float64, device=ux.device)
    _lib.check(
        ctx.lib.fava_plane_pivots(ctx.handle, _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(ux), nz, ny, nx, int(axis),
                                  _ptr(piv), _stream(ux)),
        "fava_plane_pivots",
    )
    return piv


def plane_moments(rho, ux, uy, uz, axis: int, pivots: torch.Tensor | None = None,
                  out: torch.Tensor | None = None, accumulate: bool = False):
    """Single-pass pivoted plane moments [FAVA_NMOM][nbins] (fava_plane_moments)."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    if axis not in (0, 1, 2):
        raise ValueError(f"Do not recognize AXIS enumeration {axis}")
    nbins = (nx, ny, nz)[axis]
    ctx = get_context(rho.device)
    if pivots is None:
        pivots = plane_pivots(ux, uy, uz, axis)
    if out is None:
        out = torch.empty((FAVA_NMOM, nbins), dtype=torch.float64, device=rho.device)
        accumulate = False
    _lib.check(
        ctx.lib.fava_plane_moments(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz, ny,
                                   nx, int(axis), _ptr(pivots), _ptr(out), int(bool(accumulate)), _stream(rho)),
        "fava_plane_moments",
    )
    return out, pivots


def plane_moments_xz(rho, ux, uy, uz):
    """Moments for axis x and axis z from one pass (fava_plane_moments_xz) -> ((mom_x, piv_x), (mom_z, piv_z))."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    ctx = get_context(rho.device)
    piv_x = plane_pivots(ux, uy, uz, 0)
    piv_z = plane_pivots(ux, uy, uz, 2)
    mom_x = torch.empty((FAVA_NMOM, nx), dtype=torch.float64, device=rho.device)
    mom_z = torch.empty((FAVA_NMOM, nz), dtype=torch.float64, device=rho.device)
    _lib.check(
        ctx.lib.fava_plane_moments_xz(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz, ny, nx,
                                      _ptr(piv_x), _ptr(piv_z), _ptr(mom_x), _ptr(mom_z), _stream(rho)),
        "fava_plane_moments_xz",
    )
    return (mom_x, piv_x), (mom_z, piv_z)


def plane_moments_xyz_supported(shape) -> bool:
    nz, ny, nx = (int(v) for v in shape)
    return bool(_lib.load().fava_plane_moments_xyz_supported(nz, ny, nx))


def plane_moments_xyz(rho, ux, uy, uz):
    """Moments for all three axes from ONE pass over the fields (fava_plane_moments_xyz)
    -> [(mom_x, piv_x), (mom_y, piv_y), (mom_z, piv_z)]."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    ctx = get_context(rho.device)
    piv = [plane_pivots(ux, uy, uz, ax) for ax in (0, 1, 2)]
    mom = [torch.empty((FAVA_NMOM, n), dtype=torch.float64, device=rho.device) for n in (nx, ny, nz)]
    _lib.check(
        ctx.lib.fava_plane_moments_xyz(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz, ny, nx,
                                       _ptr(piv[0]), _ptr(piv[1]), _ptr(piv[2]), _ptr(mom[0]), _ptr(mom[1]), _ptr(mom[2]),
                                       _stream(rho)),
        "fava_plane_moments_xyz",
    )
    return list(zip(mom, piv))


def moments_repivot(moments: torch.Tensor, piv_old: torch.Tensor, piv_new: torch.Tensor) -> None:
    ctx = get_context(moments.device)
    nbins = int(moments.shape[1])
    _lib.check(
        ctx.lib.fava_moments_repivot(ctx.handle, _ptr(moments), _ptr(piv_old), _ptr(piv_new), nbins,
                                     _stream(moments)),
        "fava_moments_repivot",
    )


def moments_finalize(moments: torch.Tensor, pivots: torch.Tensor, weight: float, layer_volume: float,
                     favre: bool = True) -> dict[str, torch.Tensor]:
    """Moments -> {"means":[4][N], "reynolds":[6][N], "favre_means":[3][N], "favre":[6][N]}."""
    ctx = get_context(moments.device)
    nbins = int(moments.shape[1])
    dev = moments.device
    means = torch.empty((4, nbins), dtype=torch.float64, device=dev)
    rey = torch.empty((6, nbins), dtype=torch.float64, device=dev)
    fmeans = torch.empty((3, nbins), dtype=torch.float64, device=dev) if favre else None
    fav = torch.empty((6, nbins), dtype=torch.float64, device=dev) if favre else None
    _lib.check(
        ctx.lib.fava_moments_finalize(ctx.handle, _ptr(moments), _ptr(pivots), nbins, float(weight),
                                      float(layer_volume), _ptr(means), _ptr(rey), _ptr(fmeans), _ptr(fav),
                                      _stream(moments)),
        "fava_moments_finalize",
    )
    out = {"means": means, "reynolds": rey}
    if favre:
        out["favre_means"] = fmeans
        out["favre"] = fav
    return out


def plane_profiles(rho, ux, uy, uz, axis: int, cell_volume: float, layer_volume: float, favre: bool = True):
    """Dense front end + finalize in one call (device tensors out)."""
    mom, piv = plane_moments(rho, ux, uy, uz, axis)
    return moments_finalize(mom, piv, cell_volume, layer_volume, favre=favre)


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


# ---- block-list front end (AMR / multi-block files) ------------------------------------------------
def _check_blocks(rho, ux, uy, uz) -> tuple[int, int, int, int]:
    if rho.dim() != 4:
        raise ValueError(f"expected a [block][z][y][x] array, got shape {tuple(rho.shape)}")
    for t in (rho, ux, uy, uz):
        if not t.is_cuda:
            raise ValueError("fields must be CUDA tensors (no CPU fallback)")
        if t.shape != rho.shape or t.dtype != rho.dtype or t.device != rho.device:
            raise ValueError("rho, ux, uy, uz must share shape, dtype and device")
        if not t.is_contiguous():
            raise ValueError("fields must be contiguous in [block][z][y][x] order")
    nb, nzb, nyb, nxb = (int(s) for s in rho.shape)
    return nb, nzb, nyb, nxb


LEAF_DTYPE = np.dtype([("block", "<i8"), ("ilo", "<i8"), ("scale", "<i4"), ("pad_", "<i4"), ("vol_frac", "<f8")])
PROLONG_DTYPE = np.dtype([("block", "<i8"), ("off", "<i4", (3,)), ("scale", "<i4")])


class HostTable:
    """A host array of C structs (numpy structured array) plus its ctypes pointer.  Immutable once built (the array
    is made read-only), so `uid` identifies its contents: the library reuses the device tables it derived from a
    table with the same uid without comparing the bytes again (fava_plane_moments_blocks_uid)."""

    _uids = itertools.count(1)

    def __init__(self, arr: np.ndarray, ctype):
        self.arr = np.array(arr, copy=True, order="C")  # private copy: nobody else can write to it
        self.arr.flags.writeable = False
        self.uid = next(HostTable._uids)
        self.n = int(arr.shape[0])
        self.ptr = C.cast(C.c_void_p(self.arr.ctypes.data if self.n else 0), C.POINTER(ctype))


def leaf_table(blocks, ilo, scale, vol_frac) -> HostTable:
    """fava_leaf_desc[] from host sequences (block index, first bin, lref_n, vol_frac)."""
    arr = np.zeros(len(blocks), dtype=LEAF_DTYPE)
    arr["block"], arr["ilo"], arr["scale"], arr["vol_frac"] = blocks, ilo, scale, vol_frac
    return HostTable(arr, _lib.LeafDesc)


def plane_moments_blocks(rho, ux, uy, uz, axis: int, table: HostTable, nbins: int):
    """Pivoted plane moments of the leaves in `table` -> (moments [14][nbins], pivots [3][nbins]),
    already weighted by vol_frac (fava_plane_moments_blocks)."""
    nb, nzb, nyb, nxb = _check_blocks(rho, ux, uy, uz)
    if axis not in (0, 1, 2):
        raise ValueError(f"Do not recognize AXIS enumeration {axis}")
    ctx = get_context(rho.device)
    mom = torch.empty((FAVA_NMOM, nbins), dtype=torch.float64, device=rho.device)
    piv = torch.empty((3, nbins), dtype=torch.float64, device=rho.device)
    _lib.check(
        ctx.lib.fava_plane_moments_blocks_uid(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho),
                                              nzb, nyb, nxb, int(axis), table.ptr, table.n, table.uid, int(nbins),
                                              _ptr(mom), _ptr(piv), _stream(rho)),
        "fava_plane_moments_blocks_uid",
    )
    return mom, piv


def plane_sum(field: torch.Tensor, axis: int) -> torch.Tensor:
    """Plane sums [nbins] of one dense [z][y][x] field (fava_plane_sum)."""
    nz, ny, nx = _check_fields(field, field, field, field)
    if axis not in (0, 1, 2):
        raise ValueError(f"Do not recognize AXIS enumeration {axis}")
    ctx = get_context(field.device)
    out = torch.empty((nx, ny, nz)[axis], dtype=torch.float64, device=field.device)
    _lib.check(ctx.lib.fava_plane_sum(ctx.handle, _ptr(field), _dtype_code(field), nz, ny, nx, int(axis), _ptr(out),
                                      _stream(field)), "fava_plane_sum")
    return out


def plane_sum_blocks(blocks: torch.Tensor, axis: int, table: HostTable, nbins: int) -> torch.Tensor:
    """vol_frac-weighted plane sums of the table's leaves scattered to the fine bins (fava_plane_sum_blocks)."""
    nb, nzb, nyb, nxb = _check_blocks(blocks, blocks, blocks, blocks)
    if axis not in (0, 1, 2):
        raise ValueError(f"Do not recognize AXIS enumeration {axis}")
    ctx = get_context(blocks.device)
    out = torch.empty(nbins, dtype=torch.float64, device=blocks.device)
    _lib.check(ctx.lib.fava_plane_sum_blocks_uid(ctx.handle, _ptr(blocks), _dtype_code(blocks), nzb, nyb, nxb, int(axis),
                                                 table.ptr, table.n, table.uid, int(nbins), _ptr(out),
                                                 _stream(blocks)),
               "fava_plane_sum_blocks_uid")
    return out


# ---- prolongation ------------------------------------------------------------------------------------
def prolong_table(blocks, offs_xyz, scale) -> HostTable:
    """fava_prolong_leaf[] (source block, fine-cell corner relative to the output, 2^(L-level))."""
    arr = np.zeros(len(blocks), dtype=PROLONG_DTYPE)
    arr["block"], arr["scale"] = blocks, scale
    if len(blocks):
        arr["off"] = np.asarray(offs_xyz, dtype=np.int32).reshape(-1, 3)
    return HostTable(arr, _lib.ProlongLeaf)


def prolong(blocks: torch.Tensor, table: HostTable, out_zyx: tuple[int, int, int],
            out: torch.Tensor | None = None) -> torch.Tensor:
    """Piecewise-constant injection of the table's leaves into a uniform fp64 [NZ][NY][NX] array."""
    if blocks.dim() != 4 or not blocks.is_cuda or not blocks.is_contiguous():
        raise ValueError("blocks must be a contiguous CUDA tensor [block][z][y][x]")
    _, nzb, nyb, nxb = (int(s) for s in blocks.shape)
    NZ, NY, NX = (int(v) for v in out_zyx)
    ctx = get_context(blocks.device)
    if out is None:
        out = torch.empty((NZ, NY, NX), dtype=torch.float64, device=blocks.device)
    _lib.check(
        ctx.lib.fava_prolong(ctx.handle, _ptr(blocks), _dtype_code(blocks), nzb, nyb, nxb, table.ptr, table.n, NZ, NY,
                             NX, _ptr(out), _stream(blocks)),
        "fava_prolong",
    )
    return out


# ---- kinetic-energy spectrum ---------------------------------------------------------------------------
SPECTRUM_KEYS = ("k", "total", "longitudinal", "transverse")


def ke_spectrum(rho, ux, uy, uz) -> dict[str, np.ndarray]:
    """Whole single-GPU pipeline (fava_ke_spectrum) for a cubic [N][N][N] grid -> the reference's dict."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    if not (nz == ny == nx):
        raise ValueError(f"kinetic_energy_spectra needs a cubic grid (the reference's `.T` projection, "
                         f"FlashUniform.py:281, fails otherwise); got {(nx, ny, nz)}")
    n = nx
    ctx = get_context(rho.device)
    nb = n // 2 - 1
    bufs = [np.empty(max(nb, 0), dtype=np.float64) for _ in range(4)]
    ptrs = [b.ctypes.data_as(_lib.c_double_p) for b in bufs]
    _lib.check(
        ctx.lib.fava_ke_spectrum(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), n, *ptrs,
                                 _stream(rho)),
        "fava_ke_spectrum",
    )
    return dict(zip(SPECTRUM_KEYS, bufs))


# ---- box counting / structure functions (uniform grids) -------------------------------------------------
def fractal_tiles(field: torch.Tensor, contour: float, nz: int, zf0: int, z0: int, z1: int, counts: torch.Tensor,
                  coarse: torch.Tensor) -> None:
    """Flag the contour cells of planes [z0, z1) of a [nz][ny][nx] field (`field` holds planes from zf0 on) and add
    the filled boxes of levels 0..5 to `counts`, the tile occupancy to `coarse` (fava_fractal_tiles)."""
    if field.dim() != 3 or not field.is_cuda or not field.is_contiguous():
        raise ValueError("field must be a contiguous CUDA tensor [z][y][x]")
    nzl, ny, nx = (int(v) for v in field.shape)
    ctx = get_context(field.device)
    _lib.check(ctx.lib.fava_fractal_tiles(ctx.handle, _ptr(field), _dtype_code(field), int(nz), ny, nx, int(zf0),
                                          int(zf0) + nzl, int(z0), int(z1), float(contour), _ptr(counts), _ptr(coarse),
                                          _stream(field)), "fava_fractal_tiles")


def fractal_coarse(coarse: torch.Tensor, dims_zyx, nlevels: int, counts: torch.Tensor) -> None:
    """Levels >= 6 from the tile-occupancy grid (fava_fractal_coarse)."""
    nz, ny, nx = (int(v) for v in dims_zyx)
    ctx = get_context(coarse.device)
    _lib.check(ctx.lib.fava_fractal_coarse(ctx.handle, _ptr(coarse), nz, ny, nx, int(nlevels), _ptr(counts),
                                           _stream(coarse)), "fava_fractal_coarse")


def sf_gather(points: torch.Tensor, ux, uy, uz, nz: int, zf0: int, lo, cell, err: torch.Tensor) -> torch.Tensor:
    """Velocities [npoints][3] of the cells holding `points` [npoints][3]; zeros for planes outside the held slab
    (fava_sf_gather)."""
    nzl, ny, nx = _check_fields(ux, ux, uy, uz)
    ctx = get_context(ux.device)
    npts = int(points.shape[0])
    vel = torch.empty((npts, 3), dtype=torch.float64, device=ux.device)
    h_lo = (C.c_double * 3)(*[float(v) for v in lo])
    h_cell = (C.c_double * 3)(*[float(v) for v in cell])
    _lib.check(ctx.lib.fava_sf_gather(ctx.handle, _ptr(points), npts, _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(ux),
                                      int(nz), ny, nx, int(zf0), int(zf0) + nzl, h_lo, h_cell, _ptr(vel), _ptr(err),
                                      _stream(ux)), "fava_sf_gather")
    return vel


def sf_moments(p1, p2, v1, v2, nsep: int, npoints: int, order: int, anisotropic: bool) -> torch.Tensor:
    """[2][nsep] mean longitudinal / transverse increments to the power `order` (fava_sf_moments)."""
    ctx = get_context(p1.device)
    out = torch.empty((2, nsep), dtype=torch.float64, device=p1.device)
    _lib.check(ctx.lib.fava_sf_moments(ctx.handle, _ptr(p1), _ptr(p2), _ptr(v1), _ptr(v2), int(nsep), int(npoints),
                                       int(order), int(bool(anisotropic)), _ptr(out), _stream(p1)), "fava_sf_moments")
    return out


# ---- staging ---------------------------------------------------------------------------------------------
def stage_file(path, file_offset: int, nbytes: int, out: torch.Tensor) -> torch.Tensor:
    """pread `nbytes` at `file_offset` of `path` through the pinned ring into `out` (async on the current stream)."""
    if not out.is_cuda or not out.is_contiguous() or out.numel() * out.element_size() < nbytes:
        raise ValueError("stage_file: `out` must be a contiguous CUDA tensor of at least nbytes")
    ctx = get_context(out.device)
    _lib.check(
        ctx.lib.fava_stage_h2d(ctx.handle, str(path).encode(), int(file_offset), int(nbytes), _ptr(out), _stream(out)),
        "fava_stage_h2d",
    )
    return out


def stage_host(arr: np.ndarray, out: torch.Tensor) -> torch.Tensor:
    """Host ndarray (pageable or pinned, C-contiguous) -> device tensor through the pinned ring."""
    if not arr.flags.c_contiguous:
        raise ValueError("stage_host: array must be C-contiguous")
    if not out.is_cuda or not out.is_contiguous() or out.numel() * out.element_size() != arr.nbytes:
        raise ValueError("stage_host: `out` must be a contiguous CUDA tensor of the same byte size")
    ctx = get_context(out.device)
    _lib.check(
        ctx.lib.fava_stage_host_h2d(ctx.handle, C.c_void_p(arr.ctypes.data), int(arr.nbytes), _ptr(out), _stream(out)),
        "fava_stage_host_h2d",
    )
    return out


# ---- building blocks of the slab-decomposed spectrum (raw device pointers as ints) ------------------------
def workspace(slot: int, nbytes: int, dev=None) -> int:
    """Context-owned cudaMalloc buffer (zero-filled when (re)allocated); returns the device address."""
    ctx = get_context(dev)
    p = C.c_void_p()
    _lib.check(ctx.lib.fava_workspace(ctx.handle, int(slot), int(nbytes), C.byref(p)), "fava_workspace")
    return int(p.value)


def ipc_export(ptr: int) -> bytes:
    lib = _lib.load()
    h = (C.c_ubyte * 64)()
    _lib.check(lib.fava_ipc_export(C.c_void_p(ptr), C.byref(h)), "fava_ipc_export")
    return bytes(h)


def ipc_open(handle: bytes) -> int:
    lib = _lib.load()
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _lib.check(lib.fava_ipc_open(C.byref(h), C.byref(p)), "fava_ipc_open")
    return int(p.value)


def _cur_stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def spectral_pitch(n: int) -> int:
    """Complex elements per kx row of the spectral buffers (fava_spectral_pitch): n/2 on the hand-written transform
    path (power-of-two n in [256, 2048]; the Nyquist column is not stored), n/2 + 1 on the cuFFT path."""
    return int(_lib.load().fava_spectral_pitch(int(n)))


def spectral_bytes(n: int, planes: int) -> int:
    """Bytes of one component's spectral buffer holding `planes` planes of n x pitch complex numbers."""
    return 16 * int(planes) * int(n) * spectral_pitch(n)


def ke_weight3(rho, ux, uy, uz, wx: int, wy: int, wz: int, pitch: int | None = None) -> None:
    """K4 alone (fava_ke_weight3): sqrt(rho) u_c as real rows of `pitch` doubles (default 2 (nx/2 + 1))."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    ctx = get_context(rho.device)
    pitch = 2 * (nx // 2 + 1) if pitch is None else int(pitch)
    _lib.check(ctx.lib.fava_ke_weight3(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz * ny, nx,
                                       pitch, C.c_void_p(wx), C.c_void_p(wy), C.c_void_p(wz), _stream(rho)),
               "fava_ke_weight3")


def ke_transform_x(rho, ux, uy, uz, wx: int, wy: int, wz: int) -> None:
    """Stage 1 of the transform of a z-slab [nz_local][n][n] (fava_ke_transform_x): weighting fused with the x pass
    (hand-written path) or the weighting alone (cuFFT path)."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    if ny != nx:
        raise ValueError(f"the spectrum needs square planes, got {ny} x {nx}")
    ctx = get_context(rho.device)
    _lib.check(ctx.lib.fava_ke_transform_x(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz, nx,
                                           C.c_void_p(wx), C.c_void_p(wy), C.c_void_p(wz), _stream(rho)),
               "fava_ke_transform_x")


def ke_transform_y(w: int, nz_local: int, n: int, dev) -> None:
    """Stage 2, one component, in place: complex [nz_local][ky][kx] afterwards (fava_ke_transform_y)."""
    ctx = get_context(dev)
    _lib.check(ctx.lib.fava_ke_transform_y(ctx.handle, C.c_void_p(w), int(nz_local), int(n), _cur_stream(dev)),
               "fava_ke_transform_y")


def ke_transform_z(w: int, n: int, ny_local: int, ky_of_local, dev) -> None:
    """Stage 3, one component, in place: transform along z of complex [n][ny_local][pitch] (fava_ke_transform_z)."""
    ctx = get_context(dev)
    _lib.check(ctx.lib.fava_ke_transform_z(ctx.handle, C.c_void_p(w), int(n), int(ny_local), _ptr(ky_of_local),
                                           _cur_stream(dev)), "fava_ke_transform_z")


def ke_transform_xy(rho, ux, uy, uz, wx: int, wy: int, wz: int) -> None:
    """Stages 1 + 2 of a slab (or of a chunk of planes of it) for the three components."""
    nz, _, n = (int(v) for v in rho.shape)
    ke_transform_x(rho, ux, uy, uz, wx, wy, wz)
    for w in (wx, wy, wz):
        ke_transform_y(w, nz, n, rho.device)


def fft_y_scatter(w: int, n: int, nz_chunk: int, peer_table: torch.Tensor, owner_of_ky: torch.Tensor,
                  row_of_ky: torch.Tensor, rank: int, nz_local: int, nyl: int, z_offset: int = 0, max_ctas: int = 0) -> None:
    """Stage 2 fused with the slab -> pencil exchange (fava_fft_y_scatter): y transform of nz_chunk planes whose output
    rows go straight into the owners' peer-mapped receive buffers."""
    ctx = get_context(peer_table.device)
    _lib.check(ctx.lib.fava_fft_y_scatter(ctx.handle, C.c_void_p(w), int(n), int(nz_chunk), _ptr(peer_table), _ptr(owner_of_ky),
                                          _ptr(row_of_ky), int(rank), int(nz_local), int(nyl), int(z_offset), int(max_ctas),
                                          _stream(peer_table)), "fava_fft_y_scatter")


def a2a_pack(src: int, peer_table: torch.Tensor, ky_of_dest: torch.Tensor, rank: int, world: int, nz_local: int, n: int,
             nyl: int) -> None:
    ctx = get_context(peer_table.device)
    _lib.check(ctx.lib.fava_a2a_pack(ctx.handle, C.c_void_p(src), _ptr(peer_table), _ptr(ky_of_dest), rank, world,
                                     nz_local, n, nyl, _stream(peer_table)), "fava_a2a_pack")


def spectrum_bin(fx: int, fy: int, fz: int, n: int, ny_local: int, ky_of_local, sums: torch.Tensor) -> None:
    ctx = get_context(sums.device)
    norm = 1.0 / (float(n) ** 3)
    _lib.check(ctx.lib.fava_spectrum_bin(ctx.handle, C.c_void_p(fx), C.c_void_p(fy), C.c_void_p(fz), n, ny_local,
                                         _ptr(ky_of_local), norm, _ptr(sums), _stream(sums)),
               "fava_spectrum_bin")


def spectrum_finalize(sums: torch.Tensor, n: int) -> dict[str, np.ndarray]:
    ctx = get_context(sums.device)
    nb = n // 2 - 1
    bufs = [np.empty(nb, dtype=np.float64) for _ in range(4)]
    ptrs = [b.ctypes.data_as(_lib.c_double_p) for b in bufs]
    _lib.check(ctx.lib.fava_spectrum_finalize(ctx.handle, _ptr(sums), n, *ptrs, _stream(sums)), "fava_spectrum_finalize")
    return dict(zip(SPECTRUM_KEYS, bufs))


def fft_native_supported(n: int) -> bool:
    return bool(_lib.load().fava_fft_native_supported(int(n)))


def fft_x_weight3(rho, ux, uy, uz, fx: int, fy: int, fz: int, pitch: int | None = None) -> None:
    """x pass fused with the weighting (fava_fft_x_weight3): kx = 0..nx/2-1 into complex rows of `pitch` elements."""
    nz, ny, nx = _check_fields(rho, ux, uy, uz)
    ctx = get_context(rho.device)
    pitch = nx // 2 if pitch is None else int(pitch)
    _lib.check(ctx.lib.fava_fft_x_weight3(ctx.handle, _ptr(rho), _ptr(ux), _ptr(uy), _ptr(uz), _dtype_code(rho), nz * ny, nx,
                                          pitch, C.c_void_p(fx), C.c_void_p(fy), C.c_void_p(fz), _stream(rho)),
               "fava_fft_x_weight3")


def fft_cols(data: int, n: int, pitch: int, ncols: int, d1: int, d2: int, line_dim: int, dev, prune_mode: int = 0,
             ky_of_batch=None) -> None:
    """In-place FFT of length n along dimension `line_dim` of complex [d2][d1][pitch] (fava_fft_cols)."""
    ctx = get_context(dev)
    _lib.check(ctx.lib.fava_fft_cols(ctx.handle, C.c_void_p(data), int(n), int(pitch), int(ncols), int(d1), int(d2),
                                     int(line_dim), int(prune_mode), _ptr(ky_of_batch), _cur_stream(dev)), "fava_fft_cols")
