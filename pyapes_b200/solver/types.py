"""Typed dicts of the equation DSL (reference: pyapes/solver/types.py)."""
from typing import Any, Callable, TypedDict

from torch import Tensor


class DivConfigType(TypedDict):
    limiter: str  # "none" (central) | "upwind" (reference formula) | "upwind_fd" (new)
    edge: bool


class LaplacianConfigType(TypedDict):
    edge: bool


class GradConfigType(TypedDict):
    edge: bool


class DiffFluxConfigType(TypedDict):
    edge: bool


class DdtConfigType(TypedDict):
    scheme: str


class DiscretizerConfigType(TypedDict, total=False):
    div: DivConfigType
    laplacian: LaplacianConfigType
    grad: GradConfigType
    diffFlux: DiffFluxConfigType
    ddt: DdtConfigType


# signatures of `OPStype["adjust_rhs"]` (types.py:41-42): `(var) -> Tensor` and, for Div,
# `(var_j, var_i, config) -> Tensor`
GEN_RHS = Callable[[Any], Tensor]
DIV_RHS = Callable[[Any, Any, DiscretizerConfigType], Tensor]


class OPStype(TypedDict):
    """One operator of an equation (types.py:44-70).  `A_coeffs` is a compact coefficient
    descriptor (pyapes_b200._lower.StarCoeffs / FieldCoeffs) instead of 15 full-size tensors."""

    name: str
    Aop: Callable[..., Tensor]
    target: Any
    param: tuple
    sign: float | int
    other: dict | None
    A_coeffs: Any
    adjust_rhs: Callable[..., Tensor]
