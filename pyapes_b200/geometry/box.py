"""Box geometry (reference: pyapes/geometry/box.py:12-92).  Host-only."""
from __future__ import annotations

from .basis import GeoBounder, Geometry, bound_edge_and_corner

BOX_DIM = [1, 2, 3]  # dimensions a Box may have (box.py:9)


class Box(Geometry, metaclass=GeoBounder):
    """`Box([0, 0, 0], [1, 1, 1])` or `Box[0:1, 0:1, 0:1]`; bounds are stored as floats."""

    def __init__(self, lower, upper):
        assert len(lower) == len(upper), "Box: length of inputs has to be matched!"
        self._lower = [float(v) for v in lower]
        self._upper = [float(v) for v in upper]
        self.ex, self.xp, self.face, self._dim = bound_edge_and_corner(self._lower, self._upper)
        self._config = {
            i: {"e_x": e, "x_p": x, "face": f}
            for i, (e, x, f) in enumerate(zip(self.ex, self.xp, self.face))
        }

    dim = property(lambda self: self._dim)
    type = property(lambda self: "box")
    config = property(lambda self: self._config)
    lower = property(lambda self: self._lower)
    upper = property(lambda self: self._upper)
    X = property(lambda self: self._lower[0])
    Y = property(lambda self: self._lower[1])
    Z = property(lambda self: self._lower[2])

    @property
    def size(self) -> float:
        s = 1.0
        for lo, up in zip(self._lower, self._upper):
            s *= float(up - lo)
        return s
