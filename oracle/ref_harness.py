"""Run the UNMODIFIED reference (ebrooker/FAVA at /root/reference) in this container — TEST
INFRASTRUCTURE, NOT PRODUCT CODE.

The reference imports `h5py`, `mpi4py` and `yt` at module scope (fava/mesh/FLASH/_flash.py:9-15,
fava/util/_mpi.py:4); none of them is installed here and there is no libhdf5 / MPI runtime.  Its
arithmetic is pure NumPy/SciPy, so three import shims are enough to execute it as is:

  * `h5py`   -> fava_b200.h5lite (File/Group/Dataset facade over the real on-disk bytes, so the
                reference and the GPU path parse the SAME synthetic FLASH files);
  * `mpi4py` -> a 1-rank MPI.COMM_WORLD (Allreduce = copy, allgather = [x], barrier = no-op,
                Win.Allocate_shared = a bytearray);  every reference collective is a sum / max /
                concatenation over disjoint block ranges, so one rank defines the result;
  * `yt`     -> an empty module (the yt branch is dead: USE_YT = False, _flash.py:25);
  * `builtins.Optional` is injected because FlashUniform.py:28 uses `Optional` without importing it.

`/root/reference` exists only in the build container: this module is used by
tests/golden/make_golden.py (which commits the vectors) and by CPU tests that skip when the
reference is absent.  Nothing under fava_b200/ imports it.
"""

from __future__ import annotations

import builtins
import sys
import types
import typing
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")
_REPO_ROOT = Path(__file__).resolve().parent.parent


def reference_available() -> bool:
    return (REFERENCE_ROOT / "fava" / "mesh" / "FLASH" / "_flash.py").is_file()


# ---- mpi4py shim ---------------------------------------------------------------------------------
class _Win:
    def __init__(self, size: int):
        self._buf = bytearray(int(size))

    @classmethod
    def Allocate_shared(cls, size=0, disp_unit=1, comm=None, info=None):
        return cls(size)

    def Shared_query(self, rank=0):
        return memoryview(self._buf), 1

    def Fence(self, assertion=0):
        pass

    def Free(self):
        self._buf = bytearray(0)


class _Datatype:
    def __init__(self, size: int):
        self._size = size

    def Get_size(self) -> int:
        return self._size


class _Comm:
    def Get_size(self) -> int:
        return 1

    def Get_rank(self) -> int:
        return 0

    def barrier(self):
        pass

    Barrier = barrier

    def allreduce(self, x, op=None):
        return x

    def Allreduce(self, send, recv, op=None):
        np.copyto(np.asarray(recv), np.asarray(send))

    def allgather(self, x):
        return [x]

    def bcast(self, x, root=0):
        return x

    def Abort(self, code=1):
        raise SystemExit(code)


def _make_mpi4py() -> types.ModuleType:
    pkg = types.ModuleType("mpi4py")
    mpi = types.ModuleType("mpi4py.MPI")
    mpi.COMM_WORLD = _Comm()
    mpi.Intracomm = _Comm
    mpi.Win = _Win
    mpi.Datatype = _Datatype
    mpi.buffer = memoryview
    mpi.DOUBLE = _Datatype(8)
    mpi.FLOAT = _Datatype(4)
    mpi.SUM = "SUM"
    mpi.MAX = "MAX"
    mpi.MIN = "MIN"
    pkg.MPI = mpi
    return pkg


def _make_h5py() -> types.ModuleType:
    from fava_b200 import h5lite

    mod = types.ModuleType("h5py")
    mod.File = h5lite.File
    mod.Group = h5lite.Group
    mod.Dataset = h5lite.Dataset
    mod.Datatype = type("Datatype", (), {})
    return mod


_installed = False


def install() -> None:
    """Put the shims and the reference on sys.path/sys.modules (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (it only exists in the build container)")
    if str(_REPO_ROOT) not in sys.path:
        sys.path.insert(0, str(_REPO_ROOT))
    sys.dont_write_bytecode = True  # the reference mount is read-only
    builtins.Optional = typing.Optional
    mp = _make_mpi4py()
    sys.modules.setdefault("mpi4py", mp)
    sys.modules.setdefault("mpi4py.MPI", mp.MPI)
    sys.modules.setdefault("h5py", _make_h5py())
    sys.modules.setdefault("yt", types.ModuleType("yt"))
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.append(str(REFERENCE_ROOT))
    _installed = True


def ref_modules():
    """(fava.mesh.FLASH.FLASH, fava.mesh.FLASH.FlashUniform, fava package) of the reference."""
    install()
    import logging

    logging.getLogger().setLevel(logging.CRITICAL)  # the base __init__ logs an error per construction
    import fava  # noqa: F401  (the reference package)
    from fava.mesh.FLASH import FLASH as RefAMR
    from fava.mesh.FLASH import FlashUniform as RefUniform

    return RefAMR, RefUniform, fava


# ---- the three oracle routes ---------------------------------------------------------
def ref_reynolds_stress(plt_file, raxis: int = 0):
    """fava.mesh.FLASH.FLASH(plt).load(); .reynolds_stress(raxis) — _flash.py:1506-1611."""
    RefAMR, _, _ = ref_modules()
    m = RefAMR(str(plt_file))
    m.load()
    radius, stress, means = m.reynolds_stress(raxis=raxis)
    return np.array(radius), {k: np.array(v) for k, v in stress.items()}, {k: np.array(v) for k, v in means.items()}


def ref_kinetic_energy_spectra(uniform_file):
    """FlashUniform(uniform).load(); .kinetic_energy_spectra() — FlashUniform.py:229-304."""
    _, RefUniform, _ = ref_modules()
    m = RefUniform(str(uniform_file))
    m.load()
    out = m.kinetic_energy_spectra()
    return {k: np.array(v) for k, v in out.items()}


def ref_from_amr(plt_file, subdomain_coords, refine_level=-1, fields=("dens",), filename=None):
    """FLASH(plt).load(); .from_amr(...) — _flash.py:955-1377.  Returns (mesh, {field: float64[NX,NY,NZ]})
    or (mesh, None) when the reference silently returns (subdomain outside the domain)."""
    RefAMR, _, _ = ref_modules()
    m = RefAMR(str(plt_file))
    m.load()
    before = m.nblocks
    m.from_amr(subdomain_coords=subdomain_coords, refine_level=refine_level, fields=list(fields), filename=filename)
    if m.nblocks == before and m.nblocks != 1:
        return m, None
    return m, {k: np.array(m._data[k]) for k in fields}
