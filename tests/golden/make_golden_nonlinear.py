#!/usr/bin/env python3
"""Golden fixtures for NONLINEAR advection, `fdm.div(var, var)`, from the
REAL reference (read-only at /root/reference; build container only):

    python tests/golden/make_golden_nonlinear.py      ->  tests/golden/nonlinear.pt

With a Field as `var_j` the reference rebuilds the Div coefficients from the live field on every
operator application (fdm.py:306-312); when that Field is the unknown itself the coefficients
follow the iterate of the Krylov loop (`var.set_var_tensor` rebinds it each iteration,
linalg.py:122,253).  Each case stores inputs, the operator applied to the initial field, and the
solver outcome (plus the reference's own spread under 1-ulp RHS perturbations for BiCGSTAB).
"""
from __future__ import annotations

import os
import sys
import warnings
from math import pi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (puts the shim and /root/reference on sys.path)

import torch  # noqa: E402
from pyapes.solver.fdm import FDM  # noqa: E402
from pyapes.solver.ops import Solver  # noqa: E402

warnings.filterwarnings("ignore")


def init_field(mesh, var, kind):
    x = mesh.grid
    if kind == "sin":
        v = 1.0 + 0.3 * torch.sin(2 * pi * x[0])
        for g in x[1:]:
            v = v * (1.0 + 0.2 * torch.cos(2 * pi * g))
        return v.unsqueeze(0).to(var().dtype)
    g = torch.Generator().manual_seed(5)
    return (0.5 + torch.rand(var().shape, generator=g, dtype=torch.float64)).to(var().dtype)


def rhs_field(mesh, var):
    x = mesh.grid
    v = torch.cos(2 * pi * x[0])
    for g in x[1:]:
        v = v + 0.5 * torch.sin(2 * pi * g)
    return v.unsqueeze(0).to(var().dtype)


def run(spec, limiter, nu, method, tol, max_it, init_kind, perturb=0):
    mesh, var = G.build(spec)
    init = init_field(mesh, var, init_kind)
    var.set_var_tensor(init.clone())
    rhs = rhs_field(mesh, var)
    rhs_in = rhs.clone()
    if perturb:
        gp = torch.Generator().manual_seed(perturb)
        rhs = rhs * (1 + 2.2e-16 * torch.randn(rhs.shape, generator=gp, dtype=torch.float64).to(rhs.dtype))
    fdm = FDM({"div": {"limiter": limiter, "edge": False}})
    solver = Solver({"fdm": {"method": method, "tol": tol, "max_it": max_it, "report": False}})
    solver.set_eq(fdm.div(var, var) - fdm.laplacian(nu, var) == rhs)
    aop0 = solver.Aop(var).clone()
    rhs_adj = solver.rhs.clone()
    rep = solver.solve()
    return mesh, var, init, rhs_in, rhs_adj, aop0, rep


def case(name, spec, limiter, nu, method, tol, max_it, init_kind="sin"):
    mesh, var, init, rhs, rhs_adj, aop0, rep = run(spec, limiter, nu, method, tol, max_it, init_kind)
    sol = var().clone()
    sens_itr, sens_dsol = [], 0.0
    for k in range(1, 6):
        _, v2, *_rest, rep2 = run(spec, limiter, nu, method, tol, max_it, init_kind, perturb=k)
        sens_itr.append(rep2["itr"])
        sens_dsol = max(sens_dsol, (v2() - sol).abs().max().item())
    print(f"  {name:34s} {method:9s} {limiter:7s} itr={rep['itr']:4d} tol={rep['tol']:.6e} conv={rep['converge']} "
          f"sens={sens_itr} dsol={sens_dsol:.2e} finite={bool(torch.isfinite(sol).all())}")
    return {
        "name": name, "spec": spec, "bcs": G.frozen_bcs(mesh, var), "dx": [float(d) for d in mesh._dx],
        "limiter": limiter, "nu": nu, "method": method, "tol": tol, "max_it": max_it,
        "init": init, "rhs": rhs, "rhs_adjusted": rhs_adj, "aop_init": aop0, "report": rep, "solution": sol,
        "sens_itr": sens_itr, "sens_dsol": sens_dsol,
    }


def main():
    D = lambda v: ("dirichlet", v)  # noqa: E731
    P = ("periodic", None)
    out = []
    s1d = G.dspec([0], [1], [41], [D(1.0), D(0.5)])
    s1p = G.dspec([0], [1], [41], [P, P])
    s2d = G.dspec([0, 0], [1, 1], [25, 21], [D(1.0), D(0.5), D(0.8), D(1.2)])
    s2p = G.dspec([0, 0], [1, 1], [25, 21], [P, P, D(0.8), D(1.2)])
    s3d = G.dspec([0, 0, 0], [1, 1, 1], [13, 11, 12], [D(1.0), D(0.5), D(0.8), D(1.2), D(1.0), D(0.9)])
    s2n = G.dspec([0, 0], [1, 1], [25, 21], [("neumann", 0.2), D(0.5), D(0.8), ("symmetry", None)])
    # lockstep (fixed iteration count) and converged runs
    for it in (3, 10):
        out.append(case(f"nl_1d_central_dirichlet_it{it}", s1d, "none", 0.1, "bicgstab", 1e-30, it))
        out.append(case(f"nl_1d_upwind_periodic_it{it}", s1p, "upwind", 0.1, "bicgstab", 1e-30, it))
        out.append(case(f"nl_2d_central_dirichlet_it{it}", s2d, "none", 0.1, "bicgstab", 1e-30, it))
        out.append(case(f"nl_2d_upwind_mixed_it{it}", s2n, "upwind", 0.1, "bicgstab", 1e-30, it, "rand"))
        out.append(case(f"nl_3d_central_dirichlet_it{it}", s3d, "none", 0.1, "bicgstab", 1e-30, it))
        out.append(case(f"nl_2d_central_periodic_cg_it{it}", s2p, "none", 0.1, "cg", 1e-30, it))
    out.append(case("nl_1d_central_dirichlet_conv", s1d, "none", 0.1, "bicgstab", 1e-8, 500))
    out.append(case("nl_1d_upwind_periodic_conv", s1p, "upwind", 0.1, "bicgstab", 1e-8, 500))
    torch.save(out, os.path.join(HERE, "nonlinear.pt"))
    print("nonlinear.pt", os.path.getsize(os.path.join(HERE, "nonlinear.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
