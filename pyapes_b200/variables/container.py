"""Named containers for first and second derivatives (reference: pyapes/variables/container.py).

    jac = Jac(x=..., y=...); jac.x; jac["y"]; len(jac); list(jac); jac.keys
    hess = Hess(xx=..., xy=..., yy=...); hess["yx"] is hess.xy     # keys are order-insensitive

Only the components that were given exist; asking for another raises KeyError.
"""
from __future__ import annotations

from torch import Tensor


class Derivatives:
    _names: tuple[str, ...] = ()

    def __init__(self, **components: Tensor):
        for k in components:
            if k not in self._names:
                raise TypeError(f"{type(self).__name__}: unexpected component {k!r}")
        # keep declaration order, like the reference's dataclass fields
        self.keys = [k for k in self._names if k in components]
        for k in self.keys:
            setattr(self, k, components[k])
        self.max = len(self.keys)

    def __getitem__(self, key: str) -> Tensor:
        name = "".join(sorted(key.lower()))
        if name not in self.keys:
            raise KeyError(f"Derivative: key {key} not found.")
        return getattr(self, name)

    def __getattr__(self, name: str):  # only reached for components that were not given
        if name in type(self)._names:
            raise KeyError(f"Derivative: key {name} not found.")
        raise AttributeError(name)

    def __len__(self) -> int:
        return self.max

    def __iter__(self):
        return iter([getattr(self, k) for k in self.keys])


class Jac(Derivatives):
    _names = ("x", "y", "z", "r")


class Hess(Derivatives):
    _names = ("xx", "xy", "xz", "yy", "yz", "zz", "rr", "rz")
