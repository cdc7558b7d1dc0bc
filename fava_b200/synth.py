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
               amp: float = 1.0, sigma: float = 0.25, u0: float = 0.0, dtype=np.float64, y0: int = 0,
               ny: int | None = None, x0: int = 0, nx: int | None = None) -> np.ndarray:
    """The sub-box [z0,z0+nz) x [y0,y0+ny) x [x0,x0+nx) (default: whole planes) of field `name` for a global grid
    `shape_zyx`, in file order [z][y][x]; every cell depends only on its global index, so any piece can be
    generated independently and bit-identically."""
    NZ, NY, NX = shape_zyx
    nz = NZ - z0 if nz is None else nz
    ny = NY - y0 if ny is None else ny
    nx = NX - x0 if nx is None else nx
    z = np.arange(z0, z0 + nz, dtype=np.int64)[:, None, None]
    y = np.arange(y0, y0 + ny, dtype=np.int64)[None, :, None]
    x = np.arange(x0, x0 + nx, dtype=np.int64)[None, None, :]
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
        out = np.broadcast_to(out, (nz, ny, nx))
    else:
        raise KeyError(name)
    return np.ascontiguousarray(out, dtype=dtype)


def uniform_fields(shape_zyx: tuple[int, int, int], names=FIELDS, **kw) -> dict[str, np.ndarray]:
    return {n: field_slab(n, shape_zyx, **kw) for n in names}


# ================================================================================================
# Synthetic FLASH meshes and files (written through h5lite in the exact schema the reference reads,
# SURVEY Appendix A / reference fava/mesh/FLASH/_flash.py:131-159)
# ================================================================================================
class SynthMesh:
    """Block metadata of a synthetic FLASH mesh (all arrays in file conventions, 1-based gid)."""

    def __init__(self, nb_xyz, nroot_xyz, bounds, level, origin, node_type, gid, which_child):
        self.nxb, self.nyb, self.nzb = (int(v) for v in nb_xyz)
        self.nroot = tuple(int(v) for v in nroot_xyz)  # nblockx, nblocky, nblockz
        self.bounds = np.asarray(bounds, dtype=np.float64)  # (3,2)
        self.level = np.asarray(level, dtype=np.int32)  # 1-based refinement level
        self.origin = np.asarray(origin, dtype=np.int64)  # (nb,3) block index (x,y,z) at its own level
        self.node_type = np.asarray(node_type, dtype=np.int32)
        self.gid = np.asarray(gid, dtype=np.int32)
        self.which_child = np.asarray(which_child, dtype=np.int32)

    @property
    def nblocks(self) -> int:
        return int(self.level.shape[0])

    @property
    def lmax(self) -> int:
        return int(self.level.max())

    def bbox(self, dtype=np.float64) -> np.ndarray:
        """(nblocks,3,2) block bounds; exact in fp64 for power-of-two splits of the domain."""
        ext = self.bounds[:, 1] - self.bounds[:, 0]
        nper = np.asarray(self.nroot, dtype=np.float64)[None, :] * (2.0 ** (self.level[:, None] - 1))
        lo = self.bounds[None, :, 0] + ext[None, :] * (self.origin / nper)
        hi = self.bounds[None, :, 0] + ext[None, :] * ((self.origin + 1) / nper)
        return np.stack([lo, hi], axis=-1).astype(dtype)

    def fine_dims_xyz(self) -> tuple[int, int, int]:
        s = 2 ** (self.lmax - 1)
        return (self.nroot[0] * self.nxb * s, self.nroot[1] * self.nyb * s, self.nroot[2] * self.nzb * s)


def single_block_mesh(shape_zyx, bounds=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))) -> SynthMesh:
    nz, ny, nx = shape_zyx
    gid = -np.ones((1, 15), dtype=np.int32)
    return SynthMesh((nx, ny, nz), (1, 1, 1), bounds, [1], [[0, 0, 0]], [1], gid, [-1])


def _neighbour_gid(index, level, org, nper):
    """faces -x,+x,-y,+y,-z,+z: same-level neighbour id (1-based), -1 if none, -21 at the domain boundary."""
    out = []
    for ax in range(3):
        for d in (-1, 1):
            o = list(org)
            o[ax] += d
            if o[ax] < 0 or o[ax] >= nper[ax]:
                out.append(-21)
            else:
                out.append(index.get((level, o[0], o[1], o[2]), -2) + 1)
    return out


def multiblock_mesh(nroot_xyz, nb_xyz, bounds=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))) -> SynthMesh:
    """Single-level mesh of nroot blocks (x fastest block order), every block a leaf at level 1."""
    nbx, nby, nbz = nroot_xyz
    origin = [(i, j, k) for k in range(nbz) for j in range(nby) for i in range(nbx)]
    index = {(1, *o): n for n, o in enumerate(origin)}
    gid = -np.ones((len(origin), 15), dtype=np.int32)
    for n, o in enumerate(origin):
        gid[n, :6] = _neighbour_gid(index, 1, o, nroot_xyz)
    nb = len(origin)
    return SynthMesh(nb_xyz, nroot_xyz, bounds, np.ones(nb), origin, np.ones(nb), gid, -np.ones(nb))


def octree_mesh(nroot_xyz, nb_xyz, max_level: int, *, seed: int = 7, p_refine: float = 0.45,
                bounds=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), force_path: bool = True) -> SynthMesh:
    """Octree AMR mesh in pre-order (Morton) block order, like PARAMESH writes it.  A block at level
    l < max_level is refined when a seeded hash of (l, i, j, k) falls below p_refine; with
    `force_path` the chain of first children of root block 0 is always refined so that every level
    up to max_level occurs."""
    nbx, nby, nbz = nroot_xyz
    level, origin, parent, child_no = [], [], [], []
    children: dict[int, list[int]] = {}

    def refine(l, i, j, k) -> bool:
        if l >= max_level:
            return False
        if force_path and i == 0 and j == 0 and k == 0:
            return True
        h = splitmix64(np.array([(seed * 1000003 + l) * 0x1F123BB5 + ((k * 4099 + j) * 4099 + i)], dtype=np.uint64))
        return float(_uniform01(h)[0]) < p_refine

    def visit(l, i, j, k, par, cno):
        me = len(level)
        level.append(l)
        origin.append((i, j, k))
        parent.append(par)
        child_no.append(cno)
        if refine(l, i, j, k):
            kids = []
            for c in range(8):  # child order: x fastest (PARAMESH which_child 1..8)
                ci, cj, ck = c & 1, (c >> 1) & 1, (c >> 2) & 1
                kids.append(visit(l + 1, 2 * i + ci, 2 * j + cj, 2 * k + ck, me, c + 1))
            children[me] = kids
        return me

    for k in range(nbz):
        for j in range(nby):
            for i in range(nbx):
                visit(1, i, j, k, -1, -1)
    nb = len(level)
    node_type = np.ones(nb, dtype=np.int32)
    for b, kids in children.items():
        node_type[b] = 2 if any(k not in children for k in kids) else 3
    index = {(level[n], *origin[n]): n for n in range(nb)}
    gid = -np.ones((nb, 15), dtype=np.int32)
    for n in range(nb):
        nper = [nroot_xyz[a] * 2 ** (level[n] - 1) for a in range(3)]
        nbrs = _neighbour_gid(index, level[n], origin[n], nper)
        gid[n, :6] = [v if v != -1 else -1 for v in nbrs]
        gid[n, 6] = parent[n] + 1 if parent[n] >= 0 else -1
        if n in children:
            gid[n, 7:15] = [c + 1 for c in children[n]]
    return SynthMesh(nb_xyz, nroot_xyz, bounds, level, origin, node_type, gid, child_no)


def block_fields(mesh: SynthMesh, names=("dens", "velx", "vely", "velz", "pres"), *, seed: int = 1234,
                 dtype=np.float32, **kw) -> dict[str, np.ndarray]:
    """Field data [nblocks][nzb][nyb][nxb].  Leaf cells sample the same counter-based generator as
    `field_slab`, evaluated on the block's own level grid (salted per level); parent blocks hold the
    mean of their children (restriction), as FLASH plot files do."""
    nb = mesh.nblocks
    out = {n: np.zeros((nb, mesh.nzb, mesh.nyb, mesh.nxb), dtype=np.float64) for n in names}
    kids = {b: [int(c) - 1 for c in mesh.gid[b, 7:15]] for b in range(nb) if mesh.gid[b, 7] > 0}
    for b in range(nb):
        if b in kids:
            continue
        l = int(mesh.level[b])
        s = 2 ** (l - 1)
        NZ, NY, NX = mesh.nroot[2] * mesh.nzb * s, mesh.nroot[1] * mesh.nyb * s, mesh.nroot[0] * mesh.nxb * s
        ox, oy, oz = (int(v) for v in mesh.origin[b])
        for n in names:
            out[n][b] = field_slab(n, (NZ, NY, NX), z0=oz * mesh.nzb, nz=mesh.nzb, y0=oy * mesh.nyb, ny=mesh.nyb,
                                   x0=ox * mesh.nxb, nx=mesh.nxb, seed=seed + 7919 * l, **kw)
    for b in sorted(kids, reverse=True):  # pre-order => children have larger ids; fill deepest first
        for n in names:
            acc = np.zeros((mesh.nzb, mesh.nyb, mesh.nxb))
            for c, kid in enumerate(kids[b]):
                ci, cj, ck = c & 1, (c >> 1) & 1, (c >> 2) & 1
                d = out[n][kid]
                coarse = d.reshape(mesh.nzb // 2, 2, mesh.nyb // 2, 2, mesh.nxb // 2, 2).mean(axis=(1, 3, 5))
                acc[ck * mesh.nzb // 2 : (ck + 1) * mesh.nzb // 2, cj * mesh.nyb // 2 : (cj + 1) * mesh.nyb // 2,
                    ci * mesh.nxb // 2 : (ci + 1) * mesh.nxb // 2] = coarse
            out[n][b] = acc
    return {n: v.astype(dtype) for n, v in out.items()}


def _param_rows(pairs, value_dtype, width=80):
    if value_dtype == "S":
        dt = np.dtype([("name", f"S{width}"), ("value", f"S{width}")])
        rows = [(k.ljust(width).encode(), v.ljust(width).encode()) for k, v in pairs]
    else:
        dt = np.dtype([("name", f"S{width}"), ("value", value_dtype)])
        rows = [(k.ljust(width).encode(), v) for k, v in pairs]
    return np.array(rows, dtype=dt)


def write_flash_file(path, mesh: SynthMesh, fields: dict[str, np.ndarray], *, time: float = 0.0,
                     checkpoint: bool = False, uniform3d: bool = False) -> None:
    """Write a FLASH-format HDF5 file the reference's FLASH.load / FlashUniform.load accept.

    plt files carry f32 reals, chk files f64 (the reference keys this off "chk" in the file name,
    _flash.py:80-81).  `uniform3d` writes the field datasets as 3-D [z][y][x] arrays — the layout
    FAVA's own `from_amr` -> `save` produces for a one-block mesh (_flash.py:778-799)."""
    from fava_b200 import h5lite

    real = np.float64 if checkpoint else np.float32
    nb = mesh.nblocks
    bbox = mesh.bbox(real)
    names = list(fields)
    with h5lite.File(path, "w") as f:
        f.create_dataset("integer scalars", data=_param_rows(
            [("nxb", mesh.nxb), ("nyb", mesh.nyb), ("nzb", mesh.nzb), ("dimensionality", 3), ("iprocs", 1),
             ("jprocs", 1), ("kprocs", 1), ("globalnumblocks", nb), ("nstep", 1)], "<i4"))
        f.create_dataset("real scalars", data=_param_rows([("time", float(time)), ("dt", 1e-3)], "<f8"))
        f.create_dataset("logical scalars", data=_param_rows([("corners", 0)], "<i4"))
        f.create_dataset("string scalars", data=_param_rows([("geometry", "cartesian")], "S"))
        f.create_dataset("integer runtime parameters", data=_param_rows(
            [("nblockx", mesh.nroot[0]), ("nblocky", mesh.nroot[1]), ("nblockz", mesh.nroot[2]),
             ("lrefine_max", mesh.lmax)], "<i4"))
        b = mesh.bounds
        f.create_dataset("real runtime parameters", data=_param_rows(
            [("xmin", b[0, 0]), ("xmax", b[0, 1]), ("ymin", b[1, 0]), ("ymax", b[1, 1]), ("zmin", b[2, 0]),
             ("zmax", b[2, 1])], "<f8"))
        f.create_dataset("logical runtime parameters", data=_param_rows([("restart", 0)], "<i4"))
        f.create_dataset("string runtime parameters", data=_param_rows([("geometry", "cartesian")], "S"))
        f.create_dataset("unknown names", data=np.array([[n.ljust(4).encode()] for n in names], dtype="S4"))
        f.create_dataset("bounding box", data=bbox)
        f.create_dataset("coordinates", data=bbox.mean(axis=2).astype(real))
        f.create_dataset("block size", data=(bbox[..., 1] - bbox[..., 0]).astype(real))
        f.create_dataset("refine level", data=mesh.level.astype(np.int32))
        f.create_dataset("node type", data=mesh.node_type.astype(np.int32))
        f.create_dataset("gid", data=mesh.gid.astype(np.int32))
        f.create_dataset("which child", data=mesh.which_child.astype(np.int32))
        f.create_dataset("processor number", data=np.zeros(nb, dtype=np.int32))
        f.create_dataset("bflags", data=-np.ones((nb, 1), dtype=np.int32))
        for n in names:
            arr = np.asarray(fields[n])
            if uniform3d:
                arr = arr.reshape(arr.shape[-3:])
            elif arr.ndim == 3:
                arr = arr[None, ...]
            f.create_dataset(n.ljust(4), data=arr.astype(real, copy=False))


def blocks_from_uniform(mesh: SynthMesh, arr_zyx: np.ndarray) -> np.ndarray:
    """Cut a [z][y][x] array into the blocks of a single-level `multiblock_mesh`."""
    out = np.empty((mesh.nblocks, mesh.nzb, mesh.nyb, mesh.nxb), dtype=arr_zyx.dtype)
    for b, (i, j, k) in enumerate(mesh.origin):
        out[b] = arr_zyx[k * mesh.nzb : (k + 1) * mesh.nzb, j * mesh.nyb : (j + 1) * mesh.nyb,
                         i * mesh.nxb : (i + 1) * mesh.nxb]
    return out
