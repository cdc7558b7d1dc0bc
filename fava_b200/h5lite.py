"""h5lite — the subset of HDF5 that FLASH files use, in pure Python (no libhdf5 in this image).

What FLASH (and FAVA's own writer, reference fava/mesh/FLASH/_flash.py:619-799) produce is the
classic on-disk format: superblock v0/v1, groups as symbol tables (B-tree v1 + local heap + SNOD),
version-1 object headers, and datasets with a simple dataspace, a fixed-point / float / fixed-string
/ compound datatype and a CONTIGUOUS (or compact) layout.  That is all this module reads and writes;
nested groups included (the analysis-result files of fava/model/model.py:138-185); anything else (chunked or
filtered datasets, new-style groups, variable-length types) fails loudly.

Two faces:
  * an h5py-shaped facade — `File`, `Group`, `Dataset` with `shape/dtype/nbytes`, `d[()]`,
    `d[:, "name"]`, `read_direct`, `create_dataset`, `in`, `keys()` — which is exactly the surface
    the reference touches, so the same bytes feed both the oracle and the GPU path;
  * `Dataset.extent()` -> (file_offset, nbytes): the byte range of a contiguous dataset, which is
    what the staging layer `pread`s into pinned memory (fava_stage_h2d) — no Python copy of the data.

Format reference: "HDF5 File Format Specification Version 2.0" (II.A superblock, III.A B-trees,
III.B symbol-table nodes, III.D local heaps, IV.A object headers and messages 0x1/0x3/0x5/0x8/0x11).
"""

from __future__ import annotations

import logging
import os
import struct
from pathlib import Path

import numpy as np

logger = logging.getLogger(__name__)

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
DATA_ALIGN = 4096  # raw data of large datasets starts on a page boundary (O_DIRECT-friendly)
_SMALL = 1 << 16  # datasets below this size are packed on 8-byte boundaries instead


class H5LiteError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------
# datatype message <-> numpy dtype
# ------------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _encode_dtype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.fields is not None:
        names = list(dt.names)
        body = b""
        for name in names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            nm = name.encode() + b"\0"
            nm += b"\0" * (_pad8(len(nm)) - len(nm))
            body += nm + struct.pack("<IB3xII4I", off, 0, 0, 0, 0, 0, 0, 0) + _encode_dtype(sub)
        head = struct.pack("<II", 6 | (1 << 4) | (len(names) << 8), dt.itemsize)
        return head + body
    if dt.kind == "S":
        # null-padded ASCII, like h5py's fixed-length numpy 'S' mapping
        return struct.pack("<II", 3 | (1 << 4) | (1 << 8), dt.itemsize)
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits = big | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<IIHH", 0 | (1 << 4) | (bits << 8), dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            sign, eloc, esz, msz, bias = 31, 23, 8, 23, 127
        else:
            sign, eloc, esz, msz, bias = 63, 52, 11, 52, 1023
        bits = big | 0x20 | (sign << 8)
        return struct.pack("<IIHHBBBBI", 1 | (1 << 4) | (bits << 8), dt.itemsize, 0, dt.itemsize * 8, eloc, esz, 0,
                           msz, bias)
    raise H5LiteError(f"h5lite cannot store dtype {dt}")


def _decode_dtype(buf: bytes, pos: int = 0) -> tuple[np.dtype, int]:
    """Returns (dtype, bytes consumed)."""
    word, size = struct.unpack_from("<II", buf, pos)
    cls, ver, bits = word & 0xF, (word >> 4) & 0xF, word >> 8
    p = pos + 8
    order = ">" if bits & 1 else "<"
    if cls == 0:
        _, prec = struct.unpack_from("<HH", buf, p)
        if prec != size * 8:
            raise H5LiteError("fixed-point type with padding bits is not supported")
        kind = "i" if bits & 0x08 else "u"
        return np.dtype(f"{order}{kind}{size}"), p + 4 - pos
    if cls == 1:
        if size not in (4, 8):
            raise H5LiteError(f"float type of {size} bytes is not supported")
        return np.dtype(f"{order}f{size}"), p + 12 - pos
    if cls == 3:
        return np.dtype(f"S{size}"), p - pos
    if cls == 6:
        nmemb = bits & 0xFFFF
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            end = buf.index(b"\0", p)
            name = buf[p:end].decode()
            if ver < 3:
                p += _pad8(end - p + 1)
            else:
                p = end + 1
            if ver == 3:
                nb = 1
                while (1 << (8 * nb)) <= size and nb < 4:
                    nb += 1
                off = int.from_bytes(buf[p : p + nb], "little")
                p += nb
            else:
                (off,) = struct.unpack_from("<I", buf, p)
                p += 4
            if ver == 1:
                ndims = buf[p]
                if ndims:
                    raise H5LiteError("array members of version-1 compound types are not supported")
                p += 28
            sub, used = _decode_dtype(buf, p)
            p += used
            names.append(name)
            formats.append(sub)
            offsets.append(off)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - pos
    raise H5LiteError(f"HDF5 datatype class {cls} is not supported by h5lite")


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, path: Path):
        self.path = Path(path)
        self.fd = os.open(self.path, os.O_RDONLY)
        self.size = os.fstat(self.fd).st_size
        head = self.pread(0, 128)
        if head[:8] != SIGNATURE:
            os.close(self.fd)
            raise H5LiteError(f"{path}: not an HDF5 file (no signature at offset 0)")
        ver = head[8]
        if ver > 1:
            os.close(self.fd)
            raise H5LiteError(f"{path}: superblock version {ver} (new-style file) is not supported; "
                              "FLASH writes version 0")
        self.so, self.sl = head[13], head[14]
        if self.so != 8 or self.sl != 8:
            os.close(self.fd)
            raise H5LiteError("only 8-byte offsets/lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<4Q", head, p)
        p += 32
        _, root_ohdr, cache, _ = struct.unpack_from("<QQII", head, p)
        self.root_ohdr = root_ohdr
        self.root_scratch = struct.unpack_from("<QQ", head, p + 24) if cache == 1 else None

    def close(self):
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def pread(self, off: int, n: int) -> bytes:
        return os.pread(self.fd, n, off + getattr(self, "base", 0))

    # -- object headers -----------------------------------------------------------------------
    def messages(self, addr: int) -> list[tuple[int, bytes]]:
        head = self.pread(addr, 16)
        if head[0] != 1:
            raise H5LiteError(f"object header version {head[0]} at {addr} is not supported (need version 1)")
        nmsg, _, hsize = struct.unpack_from("<HII", head, 2)
        chunks = [(addr + 16, hsize)]
        out = []
        while chunks and len(out) < nmsg:
            caddr, csize = chunks.pop(0)
            buf = self.pread(caddr, csize)
            p = 0
            while p + 8 <= len(buf) and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, p)
                body = buf[p + 8 : p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:  # continuation
                    chunks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, body))
        return out

    # -- groups ------------------------------------------------------------------------------
    def group_entries(self, btree: int, heap: int) -> dict[str, tuple[int, int, tuple]]:
        hh = self.pread(heap, 32)
        if hh[:4] != b"HEAP":
            raise H5LiteError("bad local heap signature")
        dsize, _, daddr = struct.unpack_from("<QQQ", hh, 8)
        names = self.pread(daddr, dsize)
        entries: dict[str, tuple[int, int, tuple]] = {}

        def walk(addr: int):
            node = self.pread(addr, 24)
            if node[:4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", node, 4)
                if ntype != 0:
                    raise H5LiteError("unexpected B-tree node type in a group")
                body = self.pread(addr + 24, (2 * used + 1) * 8)
                for i in range(used):
                    (child,) = struct.unpack_from("<Q", body, (2 * i + 1) * 8)
                    walk(child)
            elif node[:4] == b"SNOD":
                (nsym,) = struct.unpack_from("<H", node, 6)
                body = self.pread(addr + 8, nsym * 40)
                for i in range(nsym):
                    noff, ohdr, cache = struct.unpack_from("<QQI", body, i * 40)
                    end = names.index(b"\0", noff)
                    scratch = struct.unpack_from("<QQ", body, i * 40 + 24)
                    entries[names[noff:end].decode()] = (ohdr, cache, scratch)
            else:
                raise H5LiteError(f"bad group node signature {node[:4]!r}")

        walk(btree)
        return entries

    def group_tables(self, ohdr: int) -> tuple[int, int]:
        for mtype, body in self.messages(ohdr):
            if mtype == 0x0011:
                return struct.unpack_from("<QQ", body)
        raise H5LiteError("object is not an old-style group (no symbol-table message)")

    # -- datasets ----------------------------------------------------------------------------
    def dataset_info(self, ohdr: int):
        shape = dtype = None
        layout = None
        for mtype, body in self.messages(ohdr):
            if mtype == 0x0001:
                ver, rank, flags = body[0], body[1], body[2]
                if ver == 1:
                    p = 8
                elif ver == 2:
                    p = 4
                    if body[3] == 2:  # null dataspace
                        rank = 0
                else:
                    raise H5LiteError(f"dataspace message version {ver} is not supported")
                shape = struct.unpack_from(f"<{rank}Q", body, p) if rank else ()
            elif mtype == 0x0003:
                dtype, _ = _decode_dtype(body)
            elif mtype == 0x0008:
                ver = body[0]
                if ver == 3:
                    cls = body[1]
                    if cls == 1:
                        addr, size = struct.unpack_from("<QQ", body, 2)
                        layout = ("contiguous", addr, size)
                    elif cls == 0:
                        (size,) = struct.unpack_from("<H", body, 2)
                        layout = ("compact", body[4 : 4 + size], size)
                    else:
                        layout = ("chunked", None, None)
                elif ver in (1, 2):
                    ndim, cls = body[1], body[2]
                    p = 8
                    if cls == 1:
                        (addr,) = struct.unpack_from("<Q", body, p)
                        dims = struct.unpack_from(f"<{ndim}I", body, p + 8)
                        layout = ("contiguous", addr, int(np.prod(dims, dtype=np.int64)))
                    elif cls == 0:
                        p += 4 * ndim
                        (size,) = struct.unpack_from("<I", body, p)
                        layout = ("compact", body[p + 4 : p + 4 + size], size)
                    else:
                        layout = ("chunked", None, None)
                else:
                    raise H5LiteError(f"data layout message version {ver} is not supported")
        if shape is None or dtype is None or layout is None:
            raise H5LiteError("object is not a dataset (missing dataspace/datatype/layout)")
        return tuple(int(s) for s in shape), dtype, layout


class Dataset:
    """Read-side dataset handle (h5py.Dataset look-alike for the calls the reference makes)."""

    def __init__(self, reader: _Reader, name: str, ohdr: int):
        self._r = reader
        self.name = name
        self.shape, self.dtype, self._layout = reader.dataset_info(ohdr)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def extent(self) -> tuple[int, int]:
        """(absolute file offset, nbytes) of the raw data — contiguous datasets only."""
        kind, addr, _ = self._layout
        if kind != "contiguous":
            raise H5LiteError(f"dataset {self.name!r} has a {kind} layout; only contiguous data can be staged")
        if addr == UNDEF:
            raise H5LiteError(f"dataset {self.name!r} has no allocated storage")
        return self._r.base + addr, self.nbytes

    def _raw(self) -> np.ndarray:
        kind, a, _ = self._layout
        if kind == "compact":
            return np.frombuffer(bytes(a), dtype=self.dtype, count=self.size).reshape(self.shape).copy()
        if kind == "chunked":
            raise H5LiteError(f"dataset {self.name!r} is chunked; h5lite reads contiguous/compact layouts only")
        n = self.nbytes
        raw = np.zeros(n, dtype=np.uint8)  # bytes first: compound types with out-of-order members have no buffer view
        if a != UNDEF and n:
            off, _ = self.extent()
            view = memoryview(raw)
            done = 0
            while done < n:
                got = os.preadv(self._r.fd, [view[done : min(n, done + (1 << 30))]], off + done)
                if got <= 0:
                    raise H5LiteError(f"short read in dataset {self.name!r}")
                done += got
        return raw.view(self.dtype).reshape(self.shape)

    def read_direct(self, dest: np.ndarray) -> None:
        dest[...] = self._raw()

    def __getitem__(self, key):
        arr = self._raw()
        if isinstance(key, tuple) and key and isinstance(key[-1], str):
            field = key[-1]
            rest = key[:-1]
            sub = arr[field]
            return sub[rest if len(rest) != 1 else rest[0]]
        if isinstance(key, str):
            return arr[key]
        return arr[key]

    def __len__(self):
        return self.shape[0]


class Group:
    def __init__(self, reader: _Reader, name: str, btree: int, heap: int):
        self._r = reader
        self.name = name
        self._entries = reader.group_entries(btree, heap)

    def keys(self):
        return self._entries.keys()

    def __iter__(self):
        return iter(self._entries)

    def __len__(self):
        return len(self._entries)

    def __contains__(self, key) -> bool:
        return key in self._entries

    def __getitem__(self, key: str):
        parts = [q for q in str(key).split("/") if q]
        if len(parts) > 1:
            node = self
            for part in parts:
                node = node[part]
            return node
        if key not in self._entries:
            raise KeyError(f"Unable to open object (object {key!r} doesn't exist)")
        ohdr, cache, scratch = self._entries[key]
        if cache == 1:
            return Group(self._r, key, scratch[0], scratch[1])
        try:
            return Dataset(self._r, key, ohdr)
        except H5LiteError:
            bt, hp = self._r.group_tables(ohdr)
            return Group(self._r, key, bt, hp)


# ------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------
class _PendingDataset:
    def __init__(self, name: str, data: np.ndarray):
        self.name = name
        self.data = data
        self.shape = data.shape
        self.dtype = data.dtype

    @property
    def nbytes(self):
        return self.data.nbytes

    def __getitem__(self, key):
        if isinstance(key, tuple) and key and isinstance(key[-1], str):
            sub = self.data[key[-1]]
            rest = key[:-1]
            return sub[rest if len(rest) != 1 else rest[0]]
        return self.data[key]

    def read_direct(self, dest):
        dest[...] = self.data


def _normalise_dtype(dtype) -> np.dtype:
    """Accepts what h5py accepts at the reference's call sites: numpy dtypes / strings, list-of-tuples
    compound specs and {'names','formats','offsets','itemsize'} dicts (fava/util/_types.py:5-26)."""
    return np.dtype(dtype)


def _as_array(shape, dtype, data) -> np.ndarray:
    dt = _normalise_dtype(dtype) if dtype is not None else None
    if data is None:
        if shape is None:
            raise TypeError("One of data, shape or dtype must be specified")
        shp = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        return np.zeros(shp, dtype=dt if dt is not None else np.float32)
    arr = np.asarray(data) if dt is None else np.asarray(data, dtype=dt)
    if shape is not None:
        shp = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        if int(np.prod(shp, dtype=np.int64)) != arr.size:
            raise ValueError(f"Shape tuple is incompatible with data ({shp} vs {arr.shape})")
        arr = arr.reshape(shp)
    if arr.dtype.kind == "U":
        arr = arr.astype("S")
    if arr.dtype.kind == "O":
        raise H5LiteError("object arrays cannot be stored")
    if arr.dtype.kind == "b":
        arr = arr.astype(np.int8)
    return arr


class WGroup:
    """A group of a file open for writing ("w" or "a"): an in-memory tree written out on close.  Offers the h5py
    calls the reference's writers make: create_dataset, create_group (raises if the name exists —
    fava/model/model.py:155-158 relies on that), [], in, keys(), del."""

    def __init__(self, name: str = "/"):
        self.name = name
        self.items: dict[str, "WGroup | _PendingDataset"] = {}

    def keys(self):
        return self.items.keys()

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)

    def __contains__(self, key) -> bool:
        return key in self.items

    def __getitem__(self, key: str):
        node = self
        for part in [q for q in str(key).split("/") if q]:
            if not isinstance(node, WGroup) or part not in node.items:
                raise KeyError(f"Unable to open object (object {key!r} doesn't exist)")
            node = node.items[part]
        return node

    def __delitem__(self, key: str) -> None:
        if key not in self.items:
            raise KeyError(f"Couldn't delete link (name {key!r} doesn't exist)")
        del self.items[key]

    def create_group(self, name: str) -> "WGroup":
        if name in self.items:
            raise ValueError(f"Unable to create group (name already exists): {name!r}")
        g = WGroup(name)
        self.items[name] = g
        return g

    def create_dataset(self, name: str, shape=None, dtype=None, data=None, **_kw):
        if name in self.items:
            raise ValueError(f"Unable to create dataset (name already exists): {name!r}")
        ds = _PendingDataset(name, _as_array(shape, dtype, data))
        self.items[name] = ds
        return ds


class _Writer:
    """Serialises a WGroup tree: per group an object header (symbol-table message), one B-tree leaf-level node, a
    local heap and symbol-table nodes; per dataset an object header; then the raw data (large datasets page-aligned)."""

    LEAF_K = 4
    INTERNAL_K = 16

    def __init__(self, path: Path, root: WGroup | None = None):
        self.path = Path(path)
        self.root = root if root is not None else WGroup()

    @staticmethod
    def _dataset_header(ds: _PendingDataset, addr: int) -> bytes:
        rank = len(ds.shape)
        msgs = []
        space = struct.pack("<BBBBI", 1, rank, 0, 0, 0) + struct.pack(f"<{rank}Q", *ds.shape)
        msgs.append((0x0001, space, 0))
        msgs.append((0x0003, _encode_dtype(ds.dtype), 1))
        msgs.append((0x0005, struct.pack("<BBBB", 2, 1, 2, 0), 0))
        msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, addr if ds.nbytes else UNDEF, ds.nbytes), 0))
        body = b""
        for mtype, data, flags in msgs:
            data += b"\0" * (_pad8(len(data)) - len(data))
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body

    def flush(self):
        per = 2 * self.LEAF_K
        btree_bytes = 24 + (2 * 2 * self.INTERNAL_K + 1) * 8
        pos = 96  # superblock
        plans = []  # (group, layout dict) in allocation order
        datasets = []  # (dataset, object-header address)

        def plan_group(g: WGroup) -> dict:
            nonlocal pos
            names = sorted(g.items, key=lambda q: q.encode())
            chunks = [names[i : i + per] for i in range(0, len(names), per)] or [[]]
            if len(chunks) > 2 * self.INTERNAL_K:
                raise H5LiteError(f"too many objects in one group for h5lite ({len(names)})")
            heap = bytearray(8)  # "" at offset 0
            name_off = {}
            for n in names:
                name_off[n] = len(heap)
                b = n.encode() + b"\0"
                heap += b + b"\0" * (_pad8(len(b)) - len(b))
            lay = {"names": names, "chunks": chunks, "heap": bytes(heap), "name_off": name_off}
            lay["ohdr"] = pos
            pos += 16 + 24
            lay["btree"] = pos
            pos += btree_bytes
            lay["heap_addr"] = pos
            pos += 32
            lay["heap_data"] = pos
            pos += len(heap)
            lay["snod"] = []
            for _ in chunks:
                lay["snod"].append(pos)
                pos += 8 + per * 40
            plans.append((g, lay))
            lay["child"] = {}
            for n in names:
                node = g.items[n]
                if isinstance(node, WGroup):
                    lay["child"][n] = plan_group(node)
                else:
                    lay["child"][n] = pos
                    datasets.append((node, pos))
                    pos += len(self._dataset_header(node, 0))
            return lay

        root_lay = plan_group(self.root)
        data_addr = {}
        for ds, _ in datasets:
            nb = ds.nbytes
            align = DATA_ALIGN if nb >= _SMALL else 8
            pos = (pos + align - 1) // align * align
            data_addr[id(ds)] = pos
            pos += nb
        eof = pos

        # write next to the target and rename over it: a crash, an OOM kill or ENOSPC in the middle of a flush must
        # not destroy results that are already on disk (mode "a" re-writes the whole file, and the pipeline caches
        # its per-file results there; h5py's append never re-writes existing objects)
        tmp = self.path.with_name(f".{self.path.name}.tmp{os.getpid()}")
        try:
            self._write_file(tmp, plans, datasets, data_addr, root_lay, eof, btree_bytes, per)
            os.replace(tmp, self.path)
        finally:
            if tmp.exists():
                tmp.unlink()

    def _write_file(self, target, plans, datasets, data_addr, root_lay, eof, btree_bytes, per) -> None:
        with open(target, "wb") as f:
            sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
            sb += struct.pack("<4Q", 0, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII2Q", 0, root_lay["ohdr"], 1, 0, root_lay["btree"], root_lay["heap_addr"])
            assert len(sb) == 96
            f.write(sb)
            for g, lay in plans:
                f.seek(lay["ohdr"])
                stab = struct.pack("<HHB3x2Q", 0x0011, 16, 0, lay["btree"], lay["heap_addr"])
                f.write(struct.pack("<BBHII4x", 1, 0, 1, 1, len(stab)) + stab)
                names, chunks = lay["names"], lay["chunks"]
                used = len(chunks) if names else 0
                node = b"TREE" + struct.pack("<BBHQQ", 0, 0, used, UNDEF, UNDEF)
                keys = [0] + [lay["name_off"][c[-1]] for c in chunks if c]
                for i in range(used):
                    node += struct.pack("<QQ", keys[i], lay["snod"][i])
                node += struct.pack("<Q", keys[used] if names else 0)
                node += b"\0" * (btree_bytes - len(node))
                f.seek(lay["btree"])
                f.write(node)
                f.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(lay["heap"]), 1, lay["heap_data"]))
                f.write(lay["heap"])
                for c, addr in zip(chunks, lay["snod"]):
                    blk = b"SNOD" + struct.pack("<BBH", 1, 0, len(c))
                    for n in c:
                        child = lay["child"][n]
                        if isinstance(child, dict):  # sub-group: cache type 1, scratch = (B-tree, heap)
                            blk += struct.pack("<QQII2Q", lay["name_off"][n], child["ohdr"], 1, 0, child["btree"],
                                               child["heap_addr"])
                        else:
                            blk += struct.pack("<QQII16x", lay["name_off"][n], child, 0, 0)
                    blk += b"\0" * (8 + per * 40 - len(blk))
                    f.seek(addr)
                    f.write(blk)
            for ds, addr in datasets:
                f.seek(addr)
                f.write(self._dataset_header(ds, data_addr[id(ds)]))
            for ds, _ in datasets:
                arr = ds.data
                if arr.nbytes == 0:
                    continue
                f.seek(data_addr[id(ds)])
                flat = np.ascontiguousarray(arr)
                plain = flat.dtype.fields is None and flat.dtype.kind != "S" and flat.ndim > 0
                f.write(memoryview(flat).cast("B") if plain else flat.tobytes())
            f.truncate(eof)
            f.flush()
            os.fsync(f.fileno())


def _load_tree(group: "Group", out: WGroup) -> WGroup:
    """Read a whole file into a WGroup tree (mode "a": small result files are re-written on close)."""
    for key in group.keys():
        node = group[key]
        if isinstance(node, Group):
            _load_tree(node, out.create_group(key))
        else:
            out.items[key] = _PendingDataset(key, node[()] if node.shape else np.asarray(node._raw()))
    return out


class File:
    """h5py.File look-alike: File(name, mode) with mode "r", "w" or "a" (context manager)."""

    def __init__(self, name, mode: str = "r", **_kw):
        self.filename = str(name)
        self.mode = mode
        self._reader = None
        self._writer = None
        self._root = None
        if mode == "r":
            self._reader = _Reader(Path(name))
            self._root = self._read_root()
        elif mode in ("w", "w-", "x"):
            if mode != "w" and Path(name).exists():
                raise FileExistsError(name)
            self._writer = _Writer(Path(name))
            Path(name).touch()
            self._root = self._writer.root
        elif mode in ("a", "r+"):
            tree = WGroup()
            if Path(name).is_file() and Path(name).stat().st_size > 0:
                try:
                    self._reader = _Reader(Path(name))
                    _load_tree(self._read_root(), tree)
                    self._reader.close()
                except Exception as exc:  # a truncated / foreign file: start over instead of blocking every later run
                    logger.warning("h5lite: %s is unreadable (%s); starting from an empty file", name, exc)
                    if self._reader is not None:
                        self._reader.close()
                    tree = WGroup()
                self._reader = None
            elif mode == "r+":
                raise FileNotFoundError(name)
            self._writer = _Writer(Path(name), tree)
            self._root = tree
        else:
            raise H5LiteError(f"h5lite.File mode {mode!r} is not supported (use 'r', 'w' or 'a')")

    def _read_root(self) -> "Group":
        if self._reader.root_scratch is not None:
            bt, hp = self._reader.root_scratch
        else:
            bt, hp = self._reader.group_tables(self._reader.root_ohdr)
        return Group(self._reader, "/", bt, hp)

    # context manager / lifetime
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._writer is not None:
            self._writer.flush()
            self._writer = None
        if self._reader is not None:
            self._reader.close()
            self._reader = None

    def __del__(self):
        try:
            if self._reader is not None:
                self._reader.close()
        except Exception:
            pass

    # group protocol (delegated to the root group)
    def keys(self):
        return self._root.keys()

    def __iter__(self):
        return iter(self._root.keys())

    def __len__(self):
        return len(self._root)

    def __contains__(self, key) -> bool:
        return key in self._root

    def __getitem__(self, key: str):
        return self._root[key]

    def __delitem__(self, key: str) -> None:
        if self._writer is None:
            raise H5LiteError("file is not open for writing")
        del self._root[key]

    def create_dataset(self, name: str, shape=None, dtype=None, data=None, **kw):
        if self._writer is None:
            raise H5LiteError("file is not open for writing")
        return self._root.create_dataset(name, shape=shape, dtype=dtype, data=data, **kw)

    def create_group(self, name: str):
        if self._writer is None:
            raise H5LiteError("file is not open for writing")
        return self._root.create_group(name)


def is_hdf5(path) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(8) == SIGNATURE
    except OSError:
        return False
