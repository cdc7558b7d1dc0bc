from fava_b200.mesh.base import Mesh, Structured, Unstructured
from fava_b200.mesh.flash_mesh import FLASH
from fava_b200.mesh.flash_uniform import FlashUniform

__all__ = ["Mesh", "Structured", "Unstructured", "FLASH", "FlashUniform"]
