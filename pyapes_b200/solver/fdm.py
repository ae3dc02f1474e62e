"""Equation DSL `FDM().laplacian / .grad / .div / .ddt` (reference: pyapes/solver/fdm.py).

    solver.set_eq(fdm.laplacian(c, var) == rhs)
    solver.set_eq(fdm.div(u, var) - fdm.laplacian(nu, var) == rhs)

Semantics kept verbatim: operators are class-level singletons that collect `OPStype` dicts keyed
0..k (fdm.py:372-384); `-op` / `a - b` set `sign = -1` (fdm.py:95-105); `== rhs` stores a Tensor by
reference, a Field's tensor, or a constant broadcast to the field shape (fdm.py:75-87).
`ddt` is a stub in the reference (fdm.py:315-353, commented out of FDM); here it registers an
explicit-Euler time derivative (BASELINE.json north_star item 3).
"""
from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from pyapes_b200.solver.fdc import FDC
from pyapes_b200.solver.types import DiscretizerConfigType, OPStype
from pyapes_b200.variables import Field


class Operators:
    def __init__(self):
        self._ops: dict[int, OPStype] = {}
        self._rhs: Tensor | None = None
        self._config: DiscretizerConfigType | None = None

    @property
    def ops(self) -> dict[int, OPStype]:
        return self._ops

    @ops.setter
    def ops(self, other: dict) -> None:
        self._ops = other

    @property
    def rhs(self) -> Tensor | None:
        return self._rhs

    @rhs.setter
    def rhs(self, other: Tensor | None) -> None:
        self._rhs = other

    @property
    def var(self) -> Field:
        raise NotImplementedError

    def update_config(self, config: DiscretizerConfigType) -> None:
        self._config = config

    @property
    def config(self) -> DiscretizerConfigType | None:
        return self._config

    def __eq__(self, other):  # type: ignore[override]
        if isinstance(other, Tensor):
            self._rhs = other
        elif isinstance(other, Field):
            self._rhs = other()
        else:
            self._rhs = torch.zeros_like(self.var()) + other
            # remember that this tensor is a constant fill (valid while nobody writes to it: any
            # in-place write bumps `_version`); the explicit Euler path skips reading an all-zero RHS
            self._rhs._pa_const = (float(other), self._rhs._version)  # type: ignore[attr-defined]
        assert self._rhs.shape == self.var().shape, (
            f"FDM Operators: RHS shape {self._rhs.shape} does not match {self.var().shape}!"
        )
        return self

    __hash__ = None  # type: ignore[assignment]

    def _append(self, other: "Operators", sign: float | None) -> "Operators":
        if sign is not None:
            other.ops[0]["sign"] = sign
        self._ops[list(self._ops.keys())[-1] + 1] = other.ops[0]
        return self

    def __add__(self, other: "Operators") -> "Operators":
        return self._append(other, None)

    def __sub__(self, other: "Operators") -> "Operators":
        return self._append(other, -1)

    def __neg__(self) -> "Operators":
        self._ops[0]["sign"] = -1
        return self


def _entry(name, aop, var, param, coeffs, adjust) -> OPStype:
    return {"name": name, "Aop": aop, "target": var, "param": param, "sign": 1.0, "other": None,
            "A_coeffs": coeffs, "adjust_rhs": adjust}


class Laplacian(Operators):
    r"""`coeff * \nabla^2 var` (fdm.py:108-169); `laplacian(var)` or `laplacian(coeff, var)`."""

    def __call__(self, *inputs: Any) -> "Laplacian":
        if len(inputs) == 2:
            assert isinstance(inputs[0], (int, float, Tensor)), (
                "FDM Laplacian: if additional parameter is provided, it must be a float or Tensor!"
            )
            coeffs = float(inputs[0]) if isinstance(inputs[0], int) else inputs[0]
            var = inputs[1]
        elif len(inputs) == 1:
            coeffs, var = None, inputs[0]
        else:
            raise TypeError("FDM: invalid input type!")
        A = FDC({"laplacian": {"edge": False}}).laplacian.build_A_coeffs(var)
        self._var = var
        self._ops[0] = _entry("Laplacian", self.Aop, var, (coeffs,), A, FDC.laplacian.adjust_rhs)
        return self

    var = property(lambda self: self._var)  # type: ignore[assignment]

    @staticmethod
    def Aop(param, var: Field, A_coeffs) -> Tensor:
        out = FDC({"laplacian": {"edge": False}}).laplacian.apply(A_coeffs, var)
        return out if param is None else out * param


class Grad(Operators):
    """Central gradient (fdm.py:172-230).  Inside a solver equation it only works for 1-D meshes,
    as in the reference (ops.py:145-147)."""

    def __call__(self, *inputs: Any) -> "Grad":
        if len(inputs) == 2:
            assert isinstance(inputs[0], (float, Tensor)), (
                "FDM Grad: if additional parameter is provided, it must be a float or Tensor!"
            )
            coeffs, var = inputs
        elif len(inputs) == 1:
            assert isinstance(inputs[0], Field), "FDM Grad: invalid input type! Input must be a Field."
            coeffs, var = None, inputs[0]
        else:
            raise TypeError("FDM: invalid input type!")
        A = FDC({"grad": {"edge": False}}).grad.build_A_coeffs(var)
        self._var = var
        self._ops[0] = _entry("Grad", self.Aop, var, (coeffs,), A, FDC.grad.adjust_rhs)
        return self

    var = property(lambda self: self._var)  # type: ignore[assignment]

    @staticmethod
    def Aop(param, var: Field, A_coeffs) -> Tensor:
        out = FDC({"grad": {"edge": False}}).grad.apply(A_coeffs, var)
        return out if param is None else out * param


class Div(Operators):
    """`d(var_j var_i)/dx_j` (fdm.py:233-312): `div(var_i)`, `div(u, var_i)` with `u` a float,
    a Tensor shaped like `var_i()` or a Field.  Needs `FDM(config)` with a `"div"` key."""

    def __call__(self, *inputs: Any) -> "Div":
        if len(inputs) == 2:
            # (a Jac / Hess is an FDC-level argument only: the reference asserts here as well, fdm.py:255-259)
            assert isinstance(inputs[0], (float, Tensor, Field)), (
                "FDM Grad: if additional parameter is provided, it must be a float or Tensor or Field!"
            )
            var_j, var_i = inputs
        elif len(inputs) == 1:
            var_j, var_i = 1.0, inputs[0]
        else:
            raise TypeError("FDM: invalid input type!")
        assert isinstance(var_i, Field), "FDM Div: var_i must be a Field!"
        self._var_j, self._var_i = var_j, var_i
        assert self.config is not None, "FDM Div: config must be provided!"
        A = FDC(self.config).div.build_A_coeffs(var_j, var_i, self.config)
        self._ops[0] = _entry("Div", self.Aop, var_i, (var_j, self.config), A, FDC.div.adjust_rhs)
        return self

    var = property(lambda self: self._var_i)  # type: ignore[assignment]

    @staticmethod
    def Aop(var_j, config: DiscretizerConfigType, var_i: Field, A_coeffs) -> Tensor:
        fdc = FDC(config)
        if isinstance(var_j, Field):  # live field: the coefficients follow it (fdm.py:306-312)
            return fdc.div.apply(fdc.div.build_A_coeffs(var_j, var_i, config), var_i)
        return fdc.div.apply(A_coeffs, var_i)


def _no_adjust(var: Field) -> Tensor:
    return torch.zeros_like(var())


class Ddt(Operators):
    """Euler time derivative.  `solver.set_eq(fdm.ddt(var) + <spatial ops> == rhs)` then
    `solver.solve()` advances `var` by `n_steps` steps; the solver config's `method` picks the scheme:
      "euler":                 explicit,  var_new[slicer] = var + dt * (rhs - A_spatial(var)), BCs;
      "cg"/"bicgstab"/"jacobi": implicit,  (1/dt) var_new + A_spatial(var_new) = rhs + (1/dt) var
                                solved with that method (linalg.euler_implicit);
    each step ends with var.update_time().  Not in the reference (its Ddt registers nothing,
    fdm.py:322-339; intended semantics in tests/test_fdm.py:275-299)."""

    def __call__(self, var: Field) -> "Ddt":
        try:
            dt = var.dt
        except AttributeError:
            raise AttributeError("FDM: No time step is specified.")
        self._var = var
        self._ops[0] = _entry("Ddt", self.Aop, var, (dt,), None, _no_adjust)
        self._ops[0]["other"] = {"scheme": "euler"}
        return self

    var = property(lambda self: self._var)  # type: ignore[assignment]

    @staticmethod
    def Aop(dt: float, var: Field) -> Tensor:
        return (var() - var.VARo) / dt


class FDM:
    """Operator collection; operators are class-level singletons (fdm.py:372-391), which is why
    `Solver.set_eq` resets `eq.ops` after reading it (ops.py:79-81)."""

    laplacian: Laplacian = Laplacian()
    grad: Grad = Grad()
    div: Div = Div()
    ddt: Ddt = Ddt()

    def __init__(self, config: DiscretizerConfigType | None = None) -> None:
        if config is not None:
            self.config = config
            self.div.update_config(config)
