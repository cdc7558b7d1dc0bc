"""Device-side plumbing above the C ABI: contexts, tensor hand-off, and thin functional wrappers.

PyTorch is used only for device memory, streams and (elsewhere) torch.distributed; every number is
produced by libfava_b200's own kernels through ctypes.  All wrappers take CUDA tensors in the FLASH
file layout ([z][y][x], x fastest) and enqueue on torch's current stream.
"""

from __future__ import annotations

import ctypes as C

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
