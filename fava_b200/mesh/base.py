"""Mesh base classes (reference fava/mesh/{mesh,structured,unstructured}.py): kept so that
`Model.mesh_names()` and isinstance checks behave as in the reference."""

from abc import ABC

from fava_b200.model import Model


class Mesh(ABC):
    """Generic mesh: knows its own type name and whether a file belongs to it."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self._loaded = False

    @classmethod
    def is_this_your_mesh(cls, *args, **kwargs) -> bool:
        return False

    @property
    def mesh_type(self) -> str:
        return type(self).__name__


@Model.register_mesh()
class Structured(Mesh):
    """Structured (block / uniform grid) meshes."""


@Model.register_mesh()
class Unstructured(Mesh):
    """Unstructured data (particles); outside the grid-statistics hot path."""
