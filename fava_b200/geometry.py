"""Axis / edge / geometry enumerations used by the mesh classes (reference fava/geometry/_enums.py:4-37)."""

from enum import Enum, IntEnum


class AXIS(IntEnum):
    I = 0  # noqa: E741  (x)
    J = 1  # y
    K = 2  # z


class EDGE(Enum):
    LEFT = 1
    CENTER = 2
    RIGHT = 3


class GEOMETRY(Enum):
    CARTESIAN = "cartesian"
    CYLINDRICAL = "cylindrical"
    SPHERICAL = "spherical"
    POLAR = "polar"
