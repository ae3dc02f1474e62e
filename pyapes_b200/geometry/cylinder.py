"""Axisymmetric (r, z) geometry (reference: pyapes/geometry/cylinder.py).

Only the geometry object exists here.  The rz coefficient variants of the stencils
(tools.py:64-76,86-108) are a "next" row of SURVEY.md §8(f): a Mesh over a Cylinder raises
NotImplementedError instead of silently running something else.
"""
from __future__ import annotations

from .basis import GeoBounder, Geometry, bound_edge_and_corner


class Cylinder(Geometry, metaclass=GeoBounder):
    def __init__(self, lower, upper):
        assert len(lower) == len(upper) == 2, "Cylinder: (r, z) bounds expected"
        self._lower = [float(v) for v in lower]
        self._upper = [float(v) for v in upper]
        self.ex, self.xp, self.face, self._dim = bound_edge_and_corner(self._lower, self._upper, "rz")
        self._config = {
            i: {"e_x": e, "x_p": x, "face": f}
            for i, (e, x, f) in enumerate(zip(self.ex, self.xp, self.face))
        }

    dim = property(lambda self: self._dim)
    type = property(lambda self: "cylinder")
    config = property(lambda self: self._config)
    lower = property(lambda self: self._lower)
    upper = property(lambda self: self._upper)

    @property
    def size(self) -> float:
        from math import pi

        return pi * (self._upper[0] ** 2 - self._lower[0] ** 2) * (self._upper[1] - self._lower[1])
