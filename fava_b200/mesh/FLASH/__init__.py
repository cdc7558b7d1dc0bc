"""Import path parity with the reference: `from fava.mesh.FLASH import FLASH, FlashUniform`."""
from fava_b200.mesh.flash_mesh import FIELD_MAPPING, FLASH, MESH_MDIM, NGUARD
from fava_b200.mesh.flash_uniform import FlashUniform

__all__ = ["FLASH", "FlashUniform", "FIELD_MAPPING", "NGUARD", "MESH_MDIM"]
