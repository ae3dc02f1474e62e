"""Geometry basics (reference: pyapes/geometry/basis.py).  Host-only."""
from __future__ import annotations

from typing import Any

DIR = ["x", "y", "z"]
DIR_TO_NUM: dict[str, int] = {"x": 0, "y": 1, "z": 2}
NUM_TO_DIR: dict[int, str] = {v: k for k, v in DIR_TO_NUM.items()}
DIR_TO_NUM_RZ: dict[str, int] = {"r": 0, "z": 1}
NUM_TO_DIR_RZ: dict[int, str] = {v: k for k, v in DIR_TO_NUM_RZ.items()}
SIDE_TO_NUM: dict[str, int] = {"l": 0, "u": 1}
FDIR = ["xl", "xu", "yl", "yu", "zl", "zu"]
FDIR_RZ = ["rl", "ru", "zl", "zu"]


def n2d_coord(coord: str) -> dict[int, str]:
    if coord == "xyz":
        return NUM_TO_DIR
    if coord == "rz":
        return NUM_TO_DIR_RZ
    raise RuntimeError("DiffFlux: unknown coordinate system.")


class GeoTypeIdentifier(list):
    """`int in GeoTypeIdentifier([11, 11])` — true if any entry is an instance of the type."""

    def __contains__(self, typ: type):  # type: ignore[override]
        return any(isinstance(v, typ) for v in self)


class Geometry:
    """Interface shared by Box / Cylinder."""

    dim: int
    type: str
    size: float
    lower: list[float]
    upper: list[float]
    config: dict

    def __eq__(self, other: Any):
        return (self.lower == other.lower) and (self.size == other.size)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(lower={self.lower}, upper={self.upper}, size={self.size:.1e})"


class GeoBounder(type):
    """Slice syntax: `Box[0:1, 0:2]` == `Box([0, 0], [1, 2])` (basis.py:98-133)."""

    def __getitem__(cls, item):
        if not isinstance(item, (tuple, slice)):
            raise IndexError("GeoBounder: bounds must be a tuple of slices")
        if isinstance(item, slice):
            item = (item,)
        lower, upper = [], []
        for s in item:
            assert isinstance(s, slice)
            assert type(s.start) in (int, float) and type(s.stop) in (int, float)
            assert s.step is None, "GeoBounder: step must be None"
            lower.append(float(s.start))
            upper.append(float(s.stop))
        return cls(lower, upper)


def face_order(dim: int, coord: str = "xyz") -> list[str]:
    """Order in which the reference enumerates domain faces (basis.py:152-199): note the
    2-D order is y-faces first.  It only fixes the insertion order of Mesh.d_mask."""
    if dim == 1:
        return ["xl", "xu"]
    if dim == 2:
        return ["yl", "yu", "xl", "xu"] if coord == "xyz" else ["zl", "zu", "rl", "ru"]
    return list(FDIR)


def bound_edge_and_corner(lower: list[float], upper: list[float], coord: str = "xyz"):
    """(e_x, x_p, face, dim): per face the corner it starts at and its extent."""
    dim = len(lower)
    assert 0 < dim < 4, "Dimensions must be 1, 2 and 3!"
    assert coord in ["xyz", "rz"], "Coordinate must be either xyz or rz!"
    names = "xyz" if coord == "xyz" else "rz"
    ex, xp, faces = [], [], face_order(dim, coord)
    for f in faces:
        a = names.index(f[0])
        corner = list(lower)
        if f[1] == "u":
            corner[a] = upper[a]
        extent = [(0.0 if i == a else upper[i] - lower[i]) for i in range(dim)]
        if f[1] == "u":
            extent[a] = upper[a] - corner[a]
        xp.append(corner)
        ex.append(extent)
    return ex, xp, faces, dim
