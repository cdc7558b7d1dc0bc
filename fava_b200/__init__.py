"""fava_b200 — B200-native implementation of FAVA's grid-statistics hot path behind FAVA's Python API.

    import fava_b200 as fava
    model = fava.flash("run_dir")                 # == fava.FLASH("run_dir")
    model.load(file_index=0, file_type="plt")
    radius, stress, means = model.reynolds_stress(axis=0)
    model.mesh.from_amr(subdomain_coords=box, fields=["dens", "velx"], refine_level=-1)
    model.load(file_index=0, file_type="uni")
    spectra = model.kinetic_energy_spectra()

Importing the package does not touch the GPU; the first field access loads libfava_b200.so (ctypes) and
fails loudly if it is missing or no sm_100 device is visible — there is no CPU fallback.
"""

from fava_b200.model import FileType, Model  # noqa: F401
from fava_b200.model import FLASH  # noqa: F401  (the model class, as `fava.FLASH` in the reference)
from fava_b200 import mesh  # noqa: F401  (registers the mesh classes)
from fava_b200 import analysis  # noqa: F401  (registers the analysis methods on Model)
from fava_b200.mesh import FlashUniform, Mesh, Structured, Unstructured  # noqa: F401

__version__ = "0.1.0"


def flash(directory, name=None) -> FLASH:
    """Callable entry point promised by the reference's README (README.rst:21 `fava.flash(dir)`)."""
    return FLASH(directory, name)
