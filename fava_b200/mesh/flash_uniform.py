"""Uniform single-block FLASH dataset on the GPU — drop-in for the reference's `FlashUniform`
(fava/mesh/FLASH/FlashUniform.py) on the hot path: `load()` and `kinetic_energy_spectra()`; the plane
statistics of the parent class work here too (in the reference `FlashUniform.reynolds_stress` raises
AttributeError because `load` never assigns block ranges — SURVEY §0 item 4)."""

from __future__ import annotations

import logging

import numpy as np

from fava_b200 import dist, h5lite, spectrum, uniform_analysis
from fava_b200.mesh.flash_mesh import FLASH
from fava_b200.model import Model
from fava_b200.util import timer

logger = logging.getLogger(__name__)


@Model.register_mesh()
class FlashUniform(FLASH):
    """One uniform block (`*hdf5_uniform_NNNN`, written by `from_amr`): 3-D [z][y][x] field datasets."""

    _META_UNIFORM = ("coordinates", "block size", "bounding box", "refine level")

    @classmethod
    def is_this_your_mesh(cls, filename, *args, **kwargs) -> bool:
        return "hdf5_uniform_" in str(filename)

    def load(self) -> None:
        """The reduced metadata set of the reference (FlashUniform.py:37-83) plus what the plane
        statistics need (node type), read when present."""
        if self._filename is None or not self._filename.is_file():
            logger.error("File does not exist: %s", self._filename)
            return
        try:
            self._reset_data()
            with h5lite.File(self._filename, "r") as f:
                self._read_parameters(f)
                self._set_integers()
                self._set_reals()
                self.fields = np.squeeze(f["unknown names"][()]).astype(str).reshape(-1)
                self._read_block_metadata(f, self._META_UNIFORM)
                extra = [k for k in ("node type", "gid", "which child", "bflags") if k in f]
                self._read_block_metadata(f, extra)
                if not hasattr(self, "node_type"):
                    self.node_type = np.ones(1, dtype=np.int64)
                self._index_fields(f)
        except Exception as exc:
            logger.exception("Error reading FLASH FILE %s", self._filename)
            raise RuntimeError(f"Error reading FLASH FILE {self._filename}") from exc
        self._loaded = True

    @timer
    def kinetic_energy_spectra(self) -> dict[str, np.ndarray]:
        """{"k","total","longitudinal","transverse"}, each float64[N/2-1] (reference FlashUniform.py:229-304,
        including its transposed-operand longitudinal projection; cubic 3-D grids only)."""
        if int(self.ndim) != 3:
            raise NotImplementedError("kinetic_energy_spectra is implemented for 3-D datasets")
        dims = [int(v) for v in self.nCellsVec]
        if not (dims[0] == dims[1] == dims[2]):
            # the reference fails here too: `k[n] * ffts[n].T` cannot broadcast on a non-cubic grid
            raise ValueError(f"operands could not be broadcast together: kinetic_energy_spectra needs a cubic grid, "
                             f"got {tuple(dims)}")
        t = [self.device_data(k) for k in ("dens", "velx", "vely", "velz")]
        t = [x.reshape(x.shape[-3:]) for x in t]
        if dist.world_size() == 1:
            from fava_b200 import device

            return device.ke_spectrum(*t)
        return spectrum.slab_ke_spectrum(*t, dims[0])

    @timer
    def fractal_dimension(self, field: str, contours: list[float] | float = 0.5) -> dict:
        """Box-counting dimension of the iso-contour of `field` (reference FlashUniform.py:85-227): edge flags and
        box counts on the GPU, the log-log fit on the host."""
        return uniform_analysis.fractal_dimension(self, field, contours)

    @timer
    def structure_functions(self, num_seps: int = 100, num_points: int = 10000, sep_bounds: list[float] = [0.0, 1.0],
                            log_scale: bool = True, anistropic: bool = False) -> dict:
        """Longitudinal / transverse velocity structure functions of orders 1..10 (reference FlashUniform.py:306-445);
        point pairs from np.random like the reference, gathers and reductions on the GPU."""
        return uniform_analysis.structure_functions(self, num_seps, num_points, sep_bounds, log_scale, anistropic)
