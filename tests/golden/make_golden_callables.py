#!/usr/bin/env python3
"""Fixtures with CALLABLE boundary values (Dirichlet and Neumann, bcs.py:203-205,240-241) passed to the
reference as callables -- the tests pass the same callables to pyapes_b200 instead of frozen tensors.

    python tests/golden/make_golden_callables.py   ->  tests/golden/callables.pt
"""
from __future__ import annotations

import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import torch  # noqa: E402
from bc_callables import CALLABLES  # noqa: E402
from pyapes.solver.fdm import FDM  # noqa: E402
from pyapes.solver.linalg import _apply_bc_otf  # noqa: E402
from pyapes.solver.ops import Solver  # noqa: E402

G.CALLABLES.update(CALLABLES)


def case(name, spec, method, tol, max_it, seed=17):
    mesh, var = G.build(spec)
    g = torch.Generator().manual_seed(seed)
    rhs = (torch.rand(var().shape, generator=g, dtype=torch.float64) - 0.5).to(var().dtype)
    phi = (torch.rand(var().shape, generator=g, dtype=torch.float64) - 0.5).to(var().dtype)
    out = {}
    var.set_var_tensor(phi.clone())
    s = Solver(None)
    r0 = torch.zeros_like(var())
    s.set_eq(FDM().laplacian(1.0, var) == r0)
    out["lap"] = s.Aop(var).clone()
    out["lap_rhs_adj"] = r0.clone()
    _apply_bc_otf(var, mesh)
    out["bc_applied"] = var().clone()
    var.set_var_tensor(torch.zeros_like(var()))
    solver = Solver({"fdm": {"method": method, "tol": tol, "max_it": max_it, "report": False}})
    rhs_in = rhs.clone()
    solver.set_eq(FDM().laplacian(1.0, var) == rhs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = solver.solve()
    print(f"  {name}: {method} {rep}")
    return {"name": name, "spec": spec, "phi": phi, "rhs": rhs_in, "method": method, "tol": tol, "max_it": max_it,
            "report": rep, "solution": var().clone(), "out": out}


def main():
    spec2 = G.dspec([0, 0], [1, 1.5], [33, 34], [("neumann", "neumann_cos"), ("dirichlet", "dirichlet_sin"),
                                                  ("dirichlet", 0.25), ("symmetry", None)])
    spec2d = G.dspec([0, 0], [1, 1], [40, 36], [("dirichlet", "dirichlet_sin")] * 4)
    cases = [
        case("callable_neumann_2d_lockstep", spec2, "bicgstab", 1e-30, 20),
        case("callable_neumann_2d_converged", spec2, "bicgstab", 1e-8, 2000),
        case("callable_dirichlet_2d_cg", spec2d, "cg", 1e-8, 2000),
    ]
    torch.set_default_dtype(torch.float64)
    torch.save(cases, os.path.join(HERE, "callables.pt"))
    print(len(cases), "cases ->", os.path.getsize(os.path.join(HERE, "callables.pt")), "bytes")


if __name__ == "__main__":
    main()
