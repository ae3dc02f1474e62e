#!/usr/bin/env python3
"""Fixtures for Jac-driven advection `div(jac, var)` (fdc.py:639-664,730-735), made by
the REAL reference.  Also records what the reference does with a Hess (it raises).

    python tests/golden/make_golden_jacdiv.py   ->  tests/golden/jacdiv.pt
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import torch  # noqa: E402
from pyapes.solver.fdc import FDC, hessian, jacobian  # noqa: E402
from pyapes.solver.fdm import FDM  # noqa: E402
from pyapes.variables import Field  # noqa: E402


def case(name, spec, seed=23):
    mesh, var = G.build(spec)
    g = torch.Generator().manual_seed(seed)
    phi = (torch.rand(var().shape, generator=g, dtype=torch.float64) - 0.5).to(var().dtype)
    q = torch.rand(var().shape, generator=g, dtype=torch.float64).to(var().dtype)
    var.set_var_tensor(phi.clone())
    other = Field("q", 1, mesh, None)
    other.set_var_tensor(q.clone())
    jac, hess = jacobian(other), hessian(other)
    has_ns = any(k in ("neumann", "symmetry") for k, _ in spec["bcs"])
    out = {}
    for lim in ("upwind",) + (() if has_ns else ("none",)):
        fdc = FDC({"div": {"limiter": lim, "edge": False}})
        out[f"div_jac_{lim}"] = fdc.div(jac, var).clone()
        out[f"div_jac_{lim}_rhs_adj"] = fdc.div.rhs_adj.clone()
    errs = {}
    for lim in ("none", "upwind"):
        try:
            FDC({"div": {"limiter": lim, "edge": False}}).div(hess, var)
            errs[lim] = None
        except Exception as e:  # noqa: BLE001
            errs[lim] = type(e).__name__
    out["hess_errors"] = errs
    # (FDM().div asserts float | Tensor | Field, fdm.py:255-259: a Jac reaches Div only through FDC)
    try:
        FDM({"div": {"limiter": "upwind", "edge": False}}).div(jac, var)
        out["fdm_div_accepts_jac"] = True
    except AssertionError:
        out["fdm_div_accepts_jac"] = False
    print(f"  {name}: hess: {errs} fdm accepts jac: {out['fdm_div_accepts_jac']}")
    FDC({"laplacian": {"edge": False}, "grad": {"edge": False}, "div": {"limiter": "none", "edge": False}})
    return {"name": name, "spec": spec, "bcs": G.frozen_bcs(mesh, var), "dx": [float(d) for d in mesh._dx], "phi": phi,
            "q": q, "out": out}


def main():
    per2 = [("periodic", None), ("periodic", None), ("dirichlet", 0.0), ("dirichlet", 0.5)]
    neu3 = [("neumann", 0.4), ("dirichlet", 0.0), ("dirichlet", 0.0), ("dirichlet", 0.0), ("dirichlet", 0.0), ("neumann", -0.2)]
    cases = [
        case("jacdiv_2d_dirichlet", G.dspec([0, 0], [1, 1], [12, 10], [("dirichlet", 0.3)] * 4)),
        case("jacdiv_2d_periodic", G.dspec([0, 0], [1, 2], [11, 12], per2)),
        case("jacdiv_3d_neumann", G.dspec([0, 0, 0], [1, 1, 1], [8, 7, 10], neu3)),
        case("jacdiv_3d_dirichlet_f32", G.dspec([0, 0, 0], [1, 1, 1], [6, 9, 8], [("dirichlet", 0.1)] * 6, "single")),
    ]
    torch.set_default_dtype(torch.float64)
    torch.save(cases, os.path.join(HERE, "jacdiv.pt"))
    print(len(cases), "cases ->", os.path.getsize(os.path.join(HERE, "jacdiv.pt")), "bytes")


if __name__ == "__main__":
    main()
