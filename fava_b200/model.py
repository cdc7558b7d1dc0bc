"""Model layer: directory scan, the mesh / analysis registries (the reference's "plugin API",
fava/model/model.py:78-131) and the FLASH model that picks a mesh class per file type
(fava/model/flash.py:28-169)."""

from __future__ import annotations

from enum import Enum
from pathlib import Path

from fava_b200 import dist
from fava_b200.util import NotCallableError, timer


class Model:
    """A directory of simulation output plus registries of mesh classes and analysis methods."""

    _meshes: dict = {}
    _frontend = "Generic"

    def __init__(self, directory, name: str | None = None):
        self.directory = Path(directory)
        self.name = name
        self.mesh = None
        self.particles = None

    @property
    def directory(self) -> Path:
        return self._directory

    @directory.setter
    def directory(self, directory) -> None:
        d = Path(directory)
        if not d.is_dir():
            raise FileNotFoundError(f"Cannot find model directory: {d}")
        files = sorted(p for p in d.glob("*") if p.is_file())
        if not files:
            raise FileNotFoundError(f"The model directory is empty: {d}")
        self._directory = d
        self.files = files

    @property
    def name(self) -> str:
        return self._name

    @name.setter
    def name(self, name) -> None:
        self._name = self._directory.name if name is None else name

    def _filter_files(self, pattern: str) -> list:
        return [p for p in self.files if p.match(pattern)]

    def nfiles(self, *args, **kwargs) -> int:
        return len(self.files)

    # ---- registries -------------------------------------------------------------------------------
    @classmethod
    def register_mesh(cls):
        def decorator(mesh_cls):
            Model._meshes[mesh_cls.__name__] = mesh_cls
            return mesh_cls

        return decorator

    @classmethod
    def mesh_names(cls) -> list:
        return sorted(Model._meshes)

    @classmethod
    def register_analysis(cls, overwrite: bool = False, use_timer=None):
        """Install `func` as a method on Model under its own name — only if the name is free, unless
        `overwrite` (reference model.py:118-131)."""

        def decorator(func):
            if not callable(func):
                raise NotCallableError(func)
            if overwrite or not hasattr(cls, func.__name__):
                setattr(cls, func.__name__, timer(func) if use_timer else func)
            return func

        return decorator


    # ---- analysis-result files (reference model.py:138-193) ------------------------------------------
    def save_to_hdf5(self, data: dict, filename) -> None:
        """Write a (nested) dict of arrays / scalars to an HDF5 file: dicts become groups, leaves datasets;
        an existing file is opened in append mode and existing datasets are replaced."""
        from fava_b200 import h5lite

        path = Path(filename)
        with h5lite.File(str(path), "a" if path.is_file() else "w") as f:
            self.write_to_hdf5(f, data)

    def write_to_hdf5(self, handle, data: dict) -> None:
        import numpy as np

        for key, values in data.items():
            if isinstance(values, dict):
                try:
                    group = handle.create_group(key)
                except ValueError:  # the group exists already (append)
                    group = handle[key]
                self.write_to_hdf5(group, values)
                continue
            try:
                if key in handle.keys():
                    del handle[key]
                handle.create_dataset(key, data=np.copy(values))
            except Exception as exc:  # the reference reports and carries on (model.py:176-185)
                if dist.is_root():
                    print(exc)
                    print(f"[ERROR] in making {key} for {handle}", flush=True)

    def hdf5_key_exists(self, key: str, filename) -> bool:
        from fava_b200 import h5lite

        path = Path(filename)
        if not path.is_file():
            return False
        with h5lite.File(str(path), "r") as f:
            return key in list(f.keys())


class FileType(Enum):
    CHK = 0
    PLT = 1
    PRT = 2
    CHK_PRT = 3
    PLT_PRT = 4
    UNI = 5
    ANL = 6


_STEM = {"CHK": "chk", "PLT": "plt_cnt", "PRT": "part", "UNI": "uniform", "ANL": "analysis"}


def _as_filetype(ft) -> FileType:
    return ft if isinstance(ft, FileType) else FileType[str(ft).upper()]


class FLASH(Model):
    """FLASH output directory: `*hdf5_{chk,plt_cnt,part,uniform,analysis}_NNNN` files, indexed
    "by index" (sorted order) and "by number" (the 4-digit suffix)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        for attr, tag in (("chk_files", "chk"), ("plt_files", "plt_cnt"), ("prt_files", "part"),
                          ("uni_files", "uniform"), ("anl_files", "analysis")):
            found = self._filter_files(f"*hdf5_{tag}_????")
            setattr(self, attr, {
                "by number": {int(str(p).split(f"hdf5_{tag}_")[-1]): p for p in found},
                "by index": dict(enumerate(found)),
            })

    def _table(self, ft: FileType) -> dict:
        return {FileType.CHK: self.chk_files, FileType.PLT: self.plt_files, FileType.PRT: self.prt_files,
                FileType.UNI: self.uni_files, FileType.ANL: self.anl_files}[ft]

    def nfiles(self, *args, **kwargs) -> int:
        ft = _as_filetype(kwargs.get("file_type", FileType.CHK))
        return len(self._table(ft)["by index"])

    def load(self, file_index: int = 0, file_number: int | None = None, file_type=FileType.CHK, fields=None,
             *args, **kwargs):
        from fava_b200.mesh import FLASH as FlashAMR
        from fava_b200.mesh import FlashUniform

        ft = _as_filetype(file_type)
        fkey = "by index" if file_number is None else "by number"
        nkey = file_index if file_number is None else file_number
        self.mesh = None
        self.particles = None
        if ft in (FileType.CHK, FileType.PLT):
            table = self._table(ft)
            assert nkey in table[fkey]
            self.mesh = FlashAMR(filename=table[fkey][nkey])
            self.mesh.load(*args, **kwargs)
        elif ft is FileType.UNI:
            assert nkey in self.uni_files[fkey]
            path = self.uni_files[fkey][nkey]
            if dist.is_root():
                print(path)
            self.mesh = FlashUniform(filename=path)
            self.mesh.load(*args, **kwargs)
        else:
            raise NotImplementedError(f"{ft.name} files (particles / analysis caches) are outside the "
                                      "grid-statistics hot path of fava_b200")

    def convert_filename_type(self, current_filetype, new_filetype):
        if self.mesh is None:
            return None
        cur, new = _as_filetype(current_filetype), _as_filetype(new_filetype)
        stem = self.mesh.filename.stem.replace(_STEM[cur.name], _STEM[new.name])
        return self.mesh.filename.with_stem(stem)
