"""Deterministic synthetic FLASH fields: a counter-based generator so that any slab of
any file can be produced independently and bit-identically.

    v = g(splitmix64(seed ^ field_id*PHI ^ linear_index)),   linear_index = (z*ny + y)*nx + x  (file order)

dens = 1 + 0.5 U[0,1);  vel_c = A sin(2 pi m_c s/L) + sigma N(0,1) + U0 (large-scale shear along the
axis after the component's own, Box-Muller from two counters);  pres = 1 + U[0,1).
"""

from __future__ import annotations

import numpy as np

FIELDS = ("dens", "velx", "vely", "velz", "pres")
_FIELD_ID = {name: i + 1 for i, name in enumerate(FIELDS)}
_PHI = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output step on uint64 counters (vectorised, wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + _PHI
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform01(bits: np.ndarray) -> np.ndarray:
    """53-bit uniform in [0,1)."""
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _counters(seed: int, field: str, stream: int, lin: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (np.uint64(_FIELD_ID[field] * 4 + stream) * _PHI)
        return splitmix64(key ^ lin.astype(np.uint64))


def field_slab(name: str, shape_zyx: tuple[int, int, int], z0: int = 0, nz: int | None = None, *, seed: int = 1234,
               amp: float = 1.0, sigma: float = 0.25, u0: float = 0.0, dtype=np.float64) -> np.ndarray:
    """Planes [z0, z0+nz) of field `name` for a global grid `shape_zyx`, in file order [z][y][x]."""
    NZ, NY, NX = shape_zyx
    nz = NZ - z0 if nz is None else nz
    z = np.arange(z0, z0 + nz, dtype=np.int64)[:, None, None]
    y = np.arange(NY, dtype=np.int64)[None, :, None]
    x = np.arange(NX, dtype=np.int64)[None, None, :]
    lin = (z * NY + y) * NX + x
    if name == "dens":
        out = 1.0 + 0.5 * _uniform01(_counters(seed, name, 0, lin))
    elif name == "pres":
        out = 1.0 + _uniform01(_counters(seed, name, 0, lin))
    elif name in ("velx", "vely", "velz"):
        c = ("velx", "vely", "velz").index(name)
        u1 = _uniform01(_counters(seed, name, 0, lin))
        u2 = _uniform01(_counters(seed, name, 1, lin))
        gauss = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        # shear of component c varies along the next axis: velx(y), vely(z), velz(x)
        coord = ((y + 0.5) / NY, (z + 0.5) / NZ, (x + 0.5) / NX)[c]
        mode = (1, 2, 3)[c]
        out = amp * np.sin(2.0 * np.pi * mode * coord) + sigma * gauss + u0
        out = np.broadcast_to(out, (nz, NY, NX))
    else:
        raise KeyError(name)
    return np.ascontiguousarray(out, dtype=dtype)


def uniform_fields(shape_zyx: tuple[int, int, int], names=FIELDS, **kw) -> dict[str, np.ndarray]:
    return {n: field_slab(n, shape_zyx, **kw) for n in names}
