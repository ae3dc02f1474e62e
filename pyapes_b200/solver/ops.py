"""`Solver`: collects an equation, adjusts the RHS for the boundary conditions, and hands the
system to the CUDA solvers (reference: pyapes/solver/ops.py).

    solver = Solver({"fdm": {"method": "cg", "tol": 1e-6, "max_it": 1000, "report": True}})
    solver.set_eq(fdm.laplacian(1.0, var) == rhs)
    report = solver.solve()
"""
from __future__ import annotations

import torch
from torch import Tensor

from pyapes_b200 import _lower as L
from pyapes_b200 import _native as N
from pyapes_b200.solver.fdm import Operators
from pyapes_b200.solver.linalg import ReportType, euler_explicit, euler_implicit, solve
from pyapes_b200.solver.tools import SolverConfig
from pyapes_b200.solver.types import OPStype
from pyapes_b200.variables import Field


class Solver:
    def __init__(self, config: None | SolverConfig = None):
        self.config = config

    def set_eq(self, eq: Operators) -> None:
        """ops.py:47-81.  NOTE: like the reference, the RHS adjustment is added IN PLACE to the
        tensor the caller passed (ops.py:74,77)."""
        self.var = eq.var
        self.eqs = eq.ops
        self.rhs = eq.rhs
        # The adjustment is non-zero only next to Neumann faces (fdc.py:438,523); without any the
        # reference adds an all-zero tensor, which is skipped here (saves two full passes).
        has_neumann = any(bc.bc_type == "neumann" for bc in (self.var.bcs or []))
        if self.rhs is not None and has_neumann:
            for e in self.eqs:
                fn = self.eqs[e]["adjust_rhs"]
                if self.eqs[e]["name"] == "Div":
                    param = self.eqs[e]["param"]
                    assert len(param) == 2
                    self.rhs += fn(param[0], self.var, param[1])
                elif self.eqs[e]["name"] != "Ddt":
                    self.rhs += fn(self.var)
        eq.ops = {}
        eq.rhs = None

    def Aop(self, var: Field) -> Tensor:
        assert self.rhs is not None, "Solver: rhs is missing. Did't you forget to set equation?"
        return _Aop(var, self.eqs)

    def solve(self) -> ReportType:
        assert self.var is not None and self.rhs is not None, (
            "Solver: target variable or rhs is missing. Did't you forget to set equation?"
        )
        assert self.config is not None, "Solver: config is missing!"
        if self.eqs[0]["name"].lower() == "ddt":
            # time stepping: method "euler" = explicit Euler steps; a linear-solver name = implicit
            # Euler, one solve per step
            method = str(self.config["fdm"].get("method", "euler")).lower()
            step = euler_explicit if method == "euler" else euler_implicit
            self.report = step(self.var, self.rhs, self.eqs, self.config["fdm"], self.var.mesh)
        else:
            self.report = solve(self.var, self.rhs, _Aop, self.eqs, self.config["fdm"], self.var.mesh)
        return self.report

    def __repr__(self) -> str:
        desc = ""
        for op in self.eqs:
            desc += f"{op} - {self.eqs[op]['name']}, target: {self.eqs[op]['target']}, param: {self.eqs[op]['param']}\n"
        desc += f"{len(self.eqs)+1} - RHS, input: {self.rhs}\n"
        return desc


def _Aop(target: Field, eqs: dict[int, OPStype]) -> Tensor:
    """sum_k sign_k * Aop_k(target) over the spatial operators (ops.py:122-154) — ONE kernel
    launch for the whole sum; values on every index, wrap-around like torch.roll."""
    phi = target()
    N.require_cuda(phi, "field")
    eq, keep = L_lower_equation(eqs, target)
    grid = L.lower_grid(target.nx, target.bcs)
    out = torch.empty_like(phi)
    N.check(N.lib().pa_stencil_apply(grid, eq, N.dtype_code(phi.dtype), phi.data_ptr(), out.data_ptr(),
                                     N.current_stream(phi.device)))
    del keep
    return out


def inv_dt_of(eqs: dict[int, OPStype], dtype) -> float:
    """1/dt of the leading Ddt entry, rounded once in the field dtype."""
    return float(torch.ones(1, dtype=dtype)[0] / float(eqs[0]["param"][0]))


def L_lower_equation(eqs: dict[int, OPStype], target: Field, implicit_ddt: bool = False):
    """OPStype dicts -> pa_equation: the spatial operators in key order; with `implicit_ddt` the
    leading Ddt becomes a diagonal star operator (1/dt)*phi appended LAST, where the reference's
    `_Aop` adds the time-derivative term (ops.py:133-134,151-152)."""
    if target.dim != 1:
        raise NotImplementedError(
            "pyapes_b200: only scalar fields (Field.dim == 1) are on the CUDA path (SURVEY.md §0 item 5)"
        )
    nd, dtype = target.mesh.dim, target().dtype
    eq = N.Equation()
    keep = []
    k = 0
    for key in eqs:
        e = eqs[key]
        name = e["name"].lower()
        if name == "ddt":
            if key > 1:
                raise ValueError("FDM: ddt is not allowed in the middle of the equation!")
            continue
        if k >= N.PA_MAX_OPS:
            raise NotImplementedError(f"pyapes_b200: at most {N.PA_MAX_OPS} spatial operators per equation")
        coeffs = e["A_coeffs"]
        param = None
        nonlinear = False
        if name in ("laplacian", "grad"):
            param = e["param"][0]
            if name == "grad" and nd != 1:
                # the reference's `Ax.view(target.size)` fails for mesh.dim > 1 (ops.py:145-147)
                raise RuntimeError(
                    f"shape '{list(target.size)}' is invalid for input of size {nd * target().numel()}"
                )
        elif name == "div":
            var_j, cfg = e["param"]
            if isinstance(var_j, Field):
                from pyapes_b200.solver.fdc import FDC

                coeffs = FDC(cfg).div.build_A_coeffs(var_j, target, cfg)  # live field (fdm.py:309-312)
                nonlinear = var_j is target  # div(var, var): the coefficients follow the iterate
        op, kp = L.lower_op(coeffs, nd, dtype, sign=float(e["sign"]), param=param, field_shape=target().shape,
                            field_device=target().device)
        if name == "div" and nonlinear:
            op.adv_is_iterate = 1
        eq.ops[k] = op
        keep.append(kp)
        k += 1
    if k == 0:
        raise ValueError("pyapes_b200: equation has no spatial operator")
    if implicit_ddt and eqs[0]["name"].lower() == "ddt":
        if k >= N.PA_MAX_OPS:
            raise NotImplementedError(f"pyapes_b200: at most {N.PA_MAX_OPS - 1} spatial operators next to ddt")
        op = N.Op()
        op.kind, op.sign, op.has_param, op.param = N.OP_STAR, 1.0, 0, 1.0
        c = inv_dt_of(eqs, dtype)
        for cls in range(3):
            op.coef[2][cls][1] = c  # kernel axis 2 is active for every mesh dimension
        eq.ops[k] = op
        k += 1
    eq.nops = k
    return eq, keep
