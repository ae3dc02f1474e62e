"""Axisymmetric (r, z) geometry (reference: pyapes/geometry/cylinder.py).

`Cylinder[r0:r1, z0:z1]`: leading axis is the radius, second the axis of symmetry.  A Mesh over it
uses the rz coefficient variants of the stencils (tools.py:64-76,86-108).
"""
from __future__ import annotations

from .basis import GeoBounder, Geometry, bound_edge_and_corner


class Cylinder(Geometry, metaclass=GeoBounder):
    def __init__(self, lower, upper):
        assert len(lower) == 2 and len(upper) == 2, (
            "Cylinder: a length of inputs has to be 2 since it is axisymmetric (r-z)!)"
        )
        assert lower[0] >= 0, "Cylinder: lower bound of radius has to be larger (or equal) to 0!"
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
    X = property(lambda self: self._lower[0])
    Y = property(lambda self: self._lower[1])

    @property
    def size(self) -> float:
        from math import pi

        # the reference's formula (cylinder.py:64-74): pi (r1 - r0)^2 (z1 - z0)
        return pi * (self._upper[0] - self._lower[0]) ** 2 * (self._upper[1] - self._lower[1])
