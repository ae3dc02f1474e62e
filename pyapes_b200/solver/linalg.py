"""Matrix-free linear solvers on the GPU (reference: pyapes/solver/linalg.py).

`solve(var, rhs, Aop, eqs, config, mesh)` keeps the reference's signature, mutation rules and
report: `var` is updated (result via `var()`), `var.VARo` is the previous iterate,
`{"itr", "tol", "converge"}` is returned, `RuntimeWarning` on max-iter, `RuntimeError` on an
unknown method or a NaN/Inf tolerance, optional printed report.

The iteration itself runs in native code (csrc/api.cu run_solver): all Krylov scalars stay in a
device-side state block, convergence is latched on the device, and the host only polls that
latch every `check_every` iterations — so `itr` and `tol` are exact without a sync per iteration.
`jacobi` is new (not in the reference, linalg.py:62-69).
"""
from __future__ import annotations

import warnings
from typing import Callable, TypedDict

import torch
from torch import Tensor

from pyapes_b200 import _lower as L
from pyapes_b200 import _native as N
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.tools import FDMSolverConfig
from pyapes_b200.solver.types import OPStype
from pyapes_b200.variables import Field


class ReportType(TypedDict):
    itr: int
    tol: float
    converge: bool


_SOLVERS = {"cg": "pa_cg_solve", "bicgstab": "pa_bicgstab_solve", "jacobi": "pa_jacobi_solve"}


def solve(var: Field, rhs: Tensor, Aop: Callable, eqs: dict[int, OPStype], config: FDMSolverConfig,
          mesh: Mesh) -> ReportType:
    """Dispatch on `config["method"]` (linalg.py:33-71).  min(mesh.nx) >= 3."""
    method = config["method"]
    assert isinstance(method, str) and method is not None, "Linalg: solver method is not defined!"
    method = method.lower()
    if method == "cg":
        return cg(var, rhs, Aop, eqs, config, mesh)
    if method == "bicgstab":
        return bicgstab(var, rhs, Aop, eqs, config, mesh)
    if method == "jacobi":
        return jacobi(var, rhs, Aop, eqs, config, mesh)
    raise RuntimeError(
        f"Linalg: solver only supports CG and BICGSTAB. {method=} would be a typo or is not supported."
    )


def _run(method: str, var: Field, rhs: Tensor, eqs, config: FDMSolverConfig, mesh: Mesh,
         implicit_ddt: bool = False) -> N.Report:
    from pyapes_b200.solver.ops import L_lower_equation

    x = var()
    N.require_cuda(x, "field")
    if not isinstance(rhs, Tensor) or rhs.shape != x.shape:
        raise ValueError("Linalg: rhs must be a tensor shaped like the field")
    if rhs.device != x.device or rhs.dtype != x.dtype:
        raise ValueError("Linalg: rhs must share device and dtype with the field")
    rhs_c = rhs if rhs.is_contiguous() else rhs.contiguous()
    nd = mesh.dim
    if min(mesh.nx) < 3:
        raise ValueError("Linalg: min(mesh.nx) >= 3 is required (linalg.py:44-45)")
    code = N.dtype_code(x.dtype)
    slab = getattr(mesh, "slab", None)
    grid = L.lower_grid(mesh.nx, var.bcs, slab)
    faces, nfaces, keep_f = L.lower_faces(var.bcs, mesh.grid, x, 0, nd, frozen=True)
    eq, keep_e = L_lower_equation(eqs, var, implicit_ddt)
    lib = N.lib()
    ws_bytes = lib.pa_solver_workspace_bytes(grid, code, N.METHOD[method])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    x_alt = torch.empty_like(x)
    cfg = N.SolverCfg()
    cfg.tol = float(config["tol"])
    cfg.max_it = int(config["max_it"])
    cfg.check_every = int(config.get("check_every", 0))
    cfg.use_graph = 1 if config.get("use_graph", True) else 0
    cfg.variant = int(config.get("variant", 0))
    cfg.flags = N.FLAG_CONTRACT if config.get("contract", False) else 0
    rep = N.Report()
    if slab is not None and slab["world"] > 1:
        from pyapes_b200 import parallel

        N.check(lib.pa_solve_dist(N.METHOD[method], grid, eq, nfaces, faces, code, x.data_ptr(), x_alt.data_ptr(),
                                  rhs_c.data_ptr(), cfg, ws.data_ptr(), ws_bytes, parallel.get_comm(x.device),
                                  slab["rank"], slab["world"], rep, N.current_stream(x.device)))
    else:
        N.check(getattr(lib, _SOLVERS[method])(grid, eq, nfaces, faces, code, x.data_ptr(), x_alt.data_ptr(),
                                               rhs_c.data_ptr(), cfg, ws.data_ptr(), ws_bytes, rep,
                                               N.current_stream(x.device)))
    del keep_f, keep_e, ws
    var._last_launches = rep.launches
    if rep.swaps > 0:  # (0: the iterate was never updated, x_alt is unwritten scratch)
        if rep.result_in_alt:
            var.VAR, var.VARo = x_alt, x
        else:
            var.VARo = x_alt
    if rep.status == N.BAD_TOL:
        raise RuntimeError(f"Invalid tolerance detected! tol: {rep.tol}")
    return rep


def _finish(rep: N.Report, config: FDMSolverConfig, label: str, report_on_maxit: bool) -> ReportType:
    max_it = config["max_it"]
    if rep.status == N.MAXIT:
        warnings.warn(f"Maximum iteration reached! max_it: {max_it}", RuntimeWarning)
        if report_on_maxit and config["report"]:
            _solution_report(rep.itr, rep.tol, label)
    elif config["report"]:
        _solution_report(rep.itr, rep.tol, label)
    return _write_report(rep.itr, rep.tol, rep.itr < max_it)


def cg(var: Field, rhs: Tensor, Aop, eqs, config: FDMSolverConfig, mesh: Mesh) -> ReportType:
    """Conjugate gradient, reference recurrences and stopping rule (linalg.py:74-159):
    convergence is tested on ||x_new - x_old||_2 over the whole array; a non-converged run
    does max_it + 1 iterations."""
    return _finish(_run("cg", var, rhs, eqs, config, mesh), config, "CG", False)


def bicgstab(var: Field, rhs: Tensor, Aop, eqs, config: FDMSolverConfig, mesh: Mesh) -> ReportType:
    """BiCGSTAB (linalg.py:162-279); `tol` is the residual norm; stops at itr >= max_it."""
    return _finish(_run("bicgstab", var, rhs, eqs, config, mesh), config, "BICGSTAB", True)


def jacobi(var: Field, rhs: Tensor, Aop, eqs, config: FDMSolverConfig, mesh: Mesh) -> ReportType:
    """Point Jacobi: x_new[slicer] = x + (rhs - A x) / diag(A); BCs, stopping rule and iteration
    count conventions of `cg`.  Not in the reference."""
    return _finish(_run("jacobi", var, rhs, eqs, config, mesh), config, "JACOBI", False)


def euler_explicit(var: Field, rhs: Tensor | None, eqs, config: FDMSolverConfig, mesh: Mesh) -> ReportType:
    """`n_steps` explicit Euler steps of d(var)/dt = rhs - A_spatial(var)."""
    from pyapes_b200.solver.ops import L_lower_equation

    x = var()
    N.require_cuda(x, "field")
    nd = mesh.dim
    code = N.dtype_code(x.dtype)
    slab = getattr(mesh, "slab", None)
    grid = L.lower_grid(mesh.nx, var.bcs, slab)
    faces, nfaces, keep_f = L.lower_faces(var.bcs, mesh.grid, x, 0, nd, frozen=True)
    eq, keep_e = L_lower_equation(eqs, var)
    dt = float(eqs[0]["param"][0])
    n_steps = int(config.get("n_steps", 1))
    rhs_ptr = None
    tag = getattr(rhs, "_pa_const", None)
    if tag is not None and tag[0] == 0.0 and tag[1] == rhs._version:
        rhs = None  # `== 0.0`: a zero source is not read at all (pa_euler_steps: rhs may be NULL)
    if rhs is not None:
        rhs_c = rhs if rhs.is_contiguous() else rhs.contiguous()
        rhs_ptr = rhs_c.data_ptr()
    import ctypes as C

    alt = torch.empty_like(x)
    in_alt = C.c_int(0)
    if slab is not None and slab["world"] > 1:
        from pyapes_b200 import parallel

        N.check(N.lib().pa_euler_steps_dist(grid, eq, nfaces, faces, code, x.data_ptr(), alt.data_ptr(), rhs_ptr, dt,
                                            n_steps, C.byref(in_alt), parallel.get_comm(x.device), slab["rank"],
                                            slab["world"], N.current_stream(x.device)))
    else:
        N.check(N.lib().pa_euler_steps(grid, eq, nfaces, faces, code, x.data_ptr(), alt.data_ptr(), rhs_ptr, dt,
                                       n_steps, C.byref(in_alt), N.current_stream(x.device)))
    for _ in range(n_steps):
        var.update_time()
    if n_steps > 0:
        var.VAR, var.VARo = (alt, x) if in_alt.value else (x, alt)
    del keep_f, keep_e
    return _write_report(n_steps, 0.0, True)


def euler_implicit(var: Field, rhs: Tensor | None, eqs, config: FDMSolverConfig, mesh: Mesh) -> ReportType:
    """`n_steps` implicit Euler steps of d(var)/dt + A_spatial(var) = rhs: per step solve
        (1/dt) var_new + A_spatial(var_new) = rhs + (1/dt) var_old
    with `config["method"]` (cg / bicgstab / jacobi), initial guess var_old, then
    `var.update_time()`.  The reference's Ddt is a stub (fdm.py:315-353); this is the semantics
    its test intends (tests/test_fdm.py:275-299), SURVEY.md §8f item 4.  The report is the last
    step's; `itr` sums the steps."""
    from pyapes_b200.solver.ops import inv_dt_of

    method = str(config["method"]).lower()
    if method not in _SOLVERS:
        raise RuntimeError(
            f"Linalg: solver only supports CG and BICGSTAB. {method=} would be a typo or is not supported."
        )
    x = var()
    N.require_cuda(x, "field")
    c = inv_dt_of(eqs, x.dtype)
    n_steps = int(config.get("n_steps", 1))
    src = torch.zeros_like(x) if rhs is None else (rhs if rhs.is_contiguous() else rhs.contiguous())
    rhs_eff = torch.empty_like(x)
    label = {"cg": "CG", "bicgstab": "BICGSTAB", "jacobi": "JACOBI"}[method]
    total, rep = 0, None
    for _ in range(n_steps):
        x = var()
        N.check(N.lib().pa_axpy(N.dtype_code(x.dtype), x.numel(), c, x.data_ptr(), src.data_ptr(), rhs_eff.data_ptr(),
                                N.current_stream(x.device)))
        rep = _run(method, var, rhs_eff, eqs, config, mesh, implicit_ddt=True)
        total += rep.itr
        var.update_time()
        if rep.status == N.MAXIT:
            break
    if rep is None:
        return _write_report(0, 0.0, True)
    out = _finish(rep, config, label, method == "bicgstab")
    out["itr"] = total if n_steps > 1 else out["itr"]
    return out


def _apply_bc_otf(var: Field, mesh: Mesh) -> Field:
    """Apply every BC of `var`, faces in list order (linalg.py:282-299) — one native call."""
    if len(var.bcs) > 0:
        if mesh.obstacle is not None:
            raise NotImplementedError
        x = var()
        N.require_cuda(x, "field")
        for d in range(var.dim):
            comp = x[d]
            faces, nfaces, keep = L.lower_faces(var.bcs, mesh.grid, x, d, mesh.dim)
            grid = L.lower_grid(mesh.nx, [])
            N.check(N.lib().pa_bc_apply(grid, nfaces, faces, N.dtype_code(x.dtype), comp.data_ptr(),
                                        N.current_stream(x.device)))
            del keep
    return var


def _nan_to_num(t_in: Tensor) -> Tensor:
    return torch.nan_to_num(t_in, nan=0.0, posinf=0.0, neginf=0.0)


def _solution_report(itr: int, tol: float, method: str) -> None:
    print(f"\n{method}: The solution  converged after {itr} iteration.")
    print(f"\ttolerance: {tol}")


def _write_report(itr: int, tol: float, converge: bool) -> ReportType:
    return {"itr": itr, "tol": tol, "converge": converge}


def _tolerance_check(var_new: Tensor, var_old: Tensor) -> float:
    """max over components of ||new - old||_2; RuntimeError on NaN/Inf (linalg.py:321-338)."""
    tol = torch.stack([torch.linalg.norm(var_new[d] - var_old[d]) for d in range(var_new.shape[0])])
    if bool(torch.isnan(tol).any()) or bool(torch.isinf(tol).any()):
        raise RuntimeError(f"Invalid tolerance detected! tol: {tol}")
    return torch.max(tol).item()
