#!/usr/bin/env python3
"""Generate golden fixtures by running the REAL reference (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/ops.pt and tests/golden/solvers.pt.  Every fixture stores the inputs
(mesh spec, boundary list with callables already evaluated on their face, the field / rhs
tensors) and the reference's outputs, so that tests can replay them through the oracle and
through the CUDA path without the reference being present.
"""
from __future__ import annotations

import os
import sys
import warnings
from math import pi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from pyapes.geometry import Box  # noqa: E402
from pyapes.mesh import Mesh  # noqa: E402
from pyapes.solver.fdc import FDC  # noqa: E402
from pyapes.solver.fdm import FDM  # noqa: E402
from pyapes.solver.linalg import _apply_bc_otf  # noqa: E402
from pyapes.solver.ops import Solver  # noqa: E402
from pyapes.testing.poisson import poisson_1d_bc, poisson_2d_bc, poisson_rhs_nd  # noqa: E402
from pyapes.variables import Field  # noqa: E402
from pyapes.variables.bcs import mixed_bcs  # noqa: E402

CALLABLES = {"poisson_1d_bc": poisson_1d_bc, "poisson_2d_bc": poisson_2d_bc}
FDIR = ["xl", "xu", "yl", "yu", "zl", "zu"]


def box(lower, upper):
    return Box(list(lower), list(upper))


def build(spec):
    """spec: lower, upper, nx, dtype ("double"|"single"), bcs [(kind, value)] in FDIR order
    (rl, ru, zl, zu for spec["rz"])."""
    from pyapes.geometry import Cylinder

    geo = Cylinder(list(spec["lower"]), list(spec["upper"])) if spec.get("rz") else box(spec["lower"], spec["upper"])
    mesh = Mesh(geo, None, list(spec["nx"]), "cpu", spec["dtype"])
    vals = [CALLABLES[v] if isinstance(v, str) else v for _, v in spec["bcs"]]
    kinds = [k for k, _ in spec["bcs"]]
    cfg = mixed_bcs(vals, kinds)
    if spec.get("rz"):
        for c, f in zip(cfg, ["rl", "ru", "zl", "zu"]):
            c["bc_face"] = f
    var = Field("p", 1, mesh, {"domain": cfg, "obstacle": None})
    return mesh, var


def frozen_bcs(mesh, var):
    """[(face, kind, value)] with callables evaluated on their own face."""
    out = []
    for bc in var.bcs:
        v = bc.bc_val
        if callable(v):
            v = v(mesh.grid, bc.bc_mask, var(), bc.bc_val_opt).clone()
        out.append((bc.bc_face, bc.bc_type, v))
    return out


def rand_like(var, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(var().shape, generator=g, dtype=torch.float64).to(var().dtype)


def op_case(name, spec, seed=7, u_const=0.7):
    mesh, var = build(spec)
    phi = rand_like(var, seed) - 0.5
    var.set_var_tensor(phi.clone())
    out = {}
    dim = mesh.dim
    has_ns = any(k in ("neumann", "symmetry") for k, _ in spec["bcs"])

    # Laplacian through the solver seam (fdm.py:124-169, ops.py:122-154)
    for tag, make in (
        ("lap", lambda f: f.laplacian(var)),
        ("lap_c", lambda f: f.laplacian(0.37, var)),
        ("neg_lap_c", lambda f: -f.laplacian(2.5, var)),
    ):
        s = Solver(None)
        rhs = torch.zeros_like(var())
        s.set_eq(make(FDM()) == rhs)
        out[tag] = s.Aop(var).clone()
        out[tag + "_rhs_adj"] = rhs.clone()  # set_eq adds adjust_rhs in place (ops.py:77)

    # Grad: explicit FDC call (the only way in >1-D)
    fdc = FDC({"grad": {"edge": False}})
    out["grad"] = fdc.grad(var).clone()
    out["grad_rhs_adj"] = fdc.grad.rhs_adj.clone()

    # Div upwind with const and tensor advection (fdc.py:746-772)
    ug = torch.Generator().manual_seed(seed + 1)
    u_t = (torch.rand(var().shape, generator=ug, dtype=torch.float64) - 0.5).to(var().dtype)
    out["u_tensor"] = u_t.clone()
    cfg = {"div": {"limiter": "upwind", "edge": False}}
    fdc = FDC(cfg)
    out["div_upwind_const"] = fdc.div(u_const, var).clone()
    out["div_upwind_const_rhs_adj"] = fdc.div.rhs_adj.clone()
    out["div_upwind_tensor"] = fdc.div(u_t, var).clone()
    out["div_upwind_tensor_rhs_adj"] = fdc.div.rhs_adj.clone()
    if not has_ns:
        cfg = {"div": {"limiter": "none", "edge": False}}
        fdc = FDC(cfg)
        out["div_central_const"] = fdc.div(u_const, var).clone()
        out["div_central_tensor"] = fdc.div(u_t, var).clone()
        out["div_central_tensor_rhs_adj"] = fdc.div.rhs_adj.clone()

    # combined equation through the solver seam (adv-diff operator)
    fdm = FDM({"div": {"limiter": "upwind", "edge": False}})
    s = Solver(None)
    rhs = torch.zeros_like(var())
    s.set_eq(fdm.div(u_const, var) - fdm.laplacian(0.1, var) == rhs)
    out["advdiff"] = s.Aop(var).clone()
    out["advdiff_rhs_adj"] = rhs.clone()
    if dim == 1:
        s = Solver(None)
        rhs = torch.zeros_like(var())
        s.set_eq(FDM().grad(var) - FDM().laplacian(0.5, var) == rhs)
        out["grad_minus_lap"] = s.Aop(var).clone()
        out["grad_minus_lap_rhs_adj"] = rhs.clone()

    # BC application (linalg.py:282-299)
    var.set_var_tensor(phi.clone())
    _apply_bc_otf(var, mesh)
    out["bc_applied"] = var().clone()

    return {
        "name": name,
        "spec": spec,
        "bcs": frozen_bcs(mesh, var),
        "dx": [float(d) for d in mesh._dx],
        "phi": phi,
        "u_const": u_const,
        "out": out,
    }


def rz_bc_ru(grid, mask, *_):
    from math import cos

    return torch.exp(-grid[1][mask]) * cos(1)


def rz_bc_zl(grid, mask, *_):
    return torch.cos(grid[0][mask])


def rz_bc_zu(grid, mask, *_):
    from math import exp

    return torch.cos(grid[0][mask]) * exp(-1)


CALLABLES.update({"rz_bc_ru": rz_bc_ru, "rz_bc_zl": rz_bc_zl, "rz_bc_zu": rz_bc_zu})


def rz_op_case(name, spec, seed=21, u_const=0.4):
    """Axisymmetric operators (tools.py:64-108, fdc.py:395-448)."""
    mesh, var = build(spec)
    phi = rand_like(var, seed) - 0.5
    var.set_var_tensor(phi.clone())
    out = {}
    for tag, make in (("lap", lambda f: f.laplacian(var)), ("neg_lap_c", lambda f: -f.laplacian(1.5, var))):
        s = Solver(None)
        rhs = torch.zeros_like(var())
        s.set_eq(make(FDM()) == rhs)
        out[tag] = s.Aop(var).clone()
        out[tag + "_rhs_adj"] = rhs.clone()
    fdc = FDC({"grad": {"edge": False}})
    out["grad"] = fdc.grad(var).clone()
    fdc = FDC({"div": {"limiter": "upwind", "edge": False}})
    out["div_upwind_const"] = fdc.div(u_const, var).clone()
    if not any(k in ("neumann", "symmetry") for k, _ in spec["bcs"]):
        fdc = FDC({"div": {"limiter": "none", "edge": False}})
        out["div_central_const"] = fdc.div(u_const, var).clone()
    var.set_var_tensor(phi.clone())
    _apply_bc_otf(var, mesh)
    out["bc_applied"] = var().clone()
    return {"name": name, "spec": spec, "bcs": frozen_bcs(mesh, var), "dx": [float(d) for d in mesh._dx],
            "phi": phi, "u_const": u_const, "out": out}


def edge_case(name, spec, seed=11, u_const=0.6):
    """edge=True explicit operators + jacobian / hessian (fdc.py:203-366, 896-944)."""
    from pyapes.solver.fdc import hessian, jacobian

    mesh, var = build(spec)
    phi = rand_like(var, seed) - 0.5
    var.set_var_tensor(phi.clone())
    out = {}
    out["lap_edge"] = FDC({"laplacian": {"edge": True}}).laplacian(var).clone()
    out["grad_edge"] = FDC({"grad": {"edge": True}}).grad(var).clone()
    has_ns = any(k in ("neumann", "symmetry") for k, _ in spec["bcs"])
    if mesh.dim == 1:
        # in >1-D the reference's Div edge treatment raises IndexError (fdc.py:303 indexes var[dim])
        out["div_upwind_edge"] = FDC({"div": {"limiter": "upwind", "edge": True}}).div(u_const, var).clone()
        if not has_ns:
            out["div_central_edge"] = FDC({"div": {"limiter": "none", "edge": True}}).div(u_const, var).clone()
    else:
        try:
            FDC({"div": {"limiter": "upwind", "edge": True}}).div(u_const, var)
            out["div_edge_raises"] = False
        except IndexError:
            out["div_edge_raises"] = True
    jac = jacobian(var)
    for k in jac.keys:
        out["jac_" + k] = jac[k].clone()
    hess = hessian(var)
    for k in hess.keys:
        out["hess_" + k] = hess[k].clone()
    FDC({"laplacian": {"edge": False}, "grad": {"edge": False}, "div": {"limiter": "none", "edge": False}})
    return {"name": name, "spec": spec, "bcs": frozen_bcs(mesh, var), "dx": [float(d) for d in mesh._dx],
            "phi": phi, "u_const": u_const, "out": out}


def solver_case(name, spec, terms, rhs_kind, method, tol, max_it, init=0.0, div_cfg=None,
                keep_solution=True, _second_pass=True, _perturb=0):
    """terms: [(kind, sign, param)], rhs_kind: float | ("rand", seed) | "poisson_nd" | tensor-fn"""
    mesh, var = build(spec)
    if init != 0.0:
        var.set_var_tensor(torch.zeros_like(var()) + init)
    if isinstance(rhs_kind, tuple) and rhs_kind[0] == "rand":
        g = torch.Generator().manual_seed(rhs_kind[1])
        rhs = torch.rand(var().shape, generator=g, dtype=torch.float64).to(var().dtype)
    elif rhs_kind == "poisson_nd":
        rhs = poisson_rhs_nd(mesh, var)
    elif callable(rhs_kind):
        rhs = rhs_kind(mesh, var)
    else:
        rhs = float(rhs_kind)
    rhs_in = rhs.clone() if isinstance(rhs, torch.Tensor) else rhs
    if _perturb:
        # 1-ulp-level relative noise on the RHS: how far does the REFERENCE move against itself?
        if not isinstance(rhs, torch.Tensor):
            rhs = torch.zeros_like(var()) + rhs
        gp = torch.Generator().manual_seed(_perturb)
        rhs = rhs * (1 + 2.2e-16 * torch.randn(rhs.shape, generator=gp, dtype=torch.float64).to(rhs.dtype))

    fdm = FDM(div_cfg) if div_cfg is not None else FDM()
    eq = None
    for kind, sign, param in terms:
        if kind == "laplacian":
            op = fdm.laplacian(var) if param is None else fdm.laplacian(param, var)
        elif kind == "grad":
            op = fdm.grad(var) if param is None else fdm.grad(param, var)
        else:
            op = fdm.div(param, var)
        if eq is None:
            eq = -op if sign < 0 else op
        else:
            eq = eq - op if sign < 0 else eq + op
    solver = Solver({"fdm": {"method": method, "tol": tol, "max_it": max_it, "report": False}})
    solver.set_eq(eq == rhs)
    rhs_adjusted = solver.rhs.clone()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        report = solver.solve()
    sol = var()
    # the same solve on one thread: torch's CPU reductions change order with the thread
    # count, and the reference's BiCGSTAB is not always reproducible against itself
    if _second_pass:
        nthr = torch.get_num_threads()
        torch.set_num_threads(1)
        alt = solver_case(name, spec, terms, rhs_kind, method, tol, max_it, init, div_cfg,
                          keep_solution=False, _second_pass=False)
        torch.set_num_threads(nthr)
        report_1thr = alt["report"]
    else:
        report_1thr = None
    sens_itr, sens_dsol = None, None
    if _second_pass and method == "bicgstab" and spec["dtype"] == "double":
        sens_itr, sens_dsol = [], 0.0
        for k in range(1, 6):
            alt = solver_case(name, spec, terms, rhs_kind, method, tol, max_it, init, div_cfg,
                              keep_solution=True, _second_pass=False, _perturb=k)
            sens_itr.append(alt["report"]["itr"])
            sens_dsol = max(sens_dsol, (alt["solution"] - sol).abs().max().item())
    case = {
        "name": name,
        "spec": spec,
        "bcs": frozen_bcs(mesh, var),
        "dx": [float(d) for d in mesh._dx],
        "terms": terms,
        "div_cfg": div_cfg,
        # a seeded random rhs is regenerated by the tests instead of being stored
        "rhs": ("rand", rhs_kind[1]) if isinstance(rhs_kind, tuple) else rhs_in,
        "rhs_adjusted_sum": rhs_adjusted.double().sum().item(),
        "method": method,
        "tol": tol,
        "max_it": max_it,
        "init": init,
        "report": report,
        "report_1thr": report_1thr,
        # BiCGSTAB only: iteration counts / max solution change of the reference under five
        # 1-ulp-level RHS perturbations (its own rounding sensitivity)
        "sens_itr": sens_itr,
        "sens_dsol": sens_dsol,
        "threads": torch.get_num_threads(),
        "sol_sum": sol.double().sum().item(),
        "sol_abs_sum": sol.double().abs().sum().item(),
    }
    if keep_solution:
        case["solution"] = sol.clone()
    if _second_pass:
        print(f"  {name:34s} {method:9s} itr={report['itr']:5d}/{report_1thr['itr']:5d} (8thr/1thr) "
              f"tol={report['tol']:.16e}/{report_1thr['tol']:.3e} sum={case['sol_sum']!r} sens={sens_itr} {sens_dsol}")
    return case


def dspec(lower, upper, nx, bcs, dtype="double"):
    return {"lower": lower, "upper": upper, "nx": nx, "dtype": dtype, "bcs": bcs}


def main():
    torch.manual_seed(0)
    D0 = ("dirichlet", 0.0)
    mixed3 = [("periodic", None), ("periodic", None), ("neumann", 0.5), ("symmetry", None),
              ("dirichlet", 0.0), ("dirichlet", 0.0)]
    mixed3b = [("neumann", -0.3), ("dirichlet", 1.5), ("symmetry", None), ("neumann", 0.25),
               ("periodic", None), ("periodic", None)]
    ops = []
    print("operator fixtures")
    for dt in ("double", "single"):
        tag = "f64" if dt == "double" else "f32"
        ops += [
            op_case(f"3d_dirichlet_{tag}", dspec([0, 0, 0], [1, 1, 1], [7, 6, 9], [("dirichlet", 0.3)] * 6, dt)),
            op_case(f"3d_mixed_{tag}", dspec([0, 0, 0], [1, 2, 1.5], [8, 7, 10], mixed3, dt)),
            op_case(f"3d_mixed_b_{tag}", dspec([-1, 0, 0], [1, 1, 1], [6, 9, 8], mixed3b, dt)),
            op_case(f"3d_tiny_n3_{tag}", dspec([0, 0, 0], [1, 1, 1], [3, 4, 3],
                                              [("neumann", 1.0), ("symmetry", None), ("symmetry", None),
                                               ("neumann", -1.0), ("neumann", 0.5), ("neumann", 0.5)], dt)),
            op_case(f"2d_dirichlet_fn_{tag}", dspec([0, 0], [1, 1], [9, 8], [("dirichlet", "poisson_2d_bc")] * 4, dt)),
            op_case(f"2d_mixed_{tag}", dspec([0, 0], [1, 0.5], [10, 12],
                                            [("neumann", 0.0), ("dirichlet", 0.0), ("neumann", 1.0), ("dirichlet", 1.0)], dt)),
            op_case(f"2d_periodic_{tag}", dspec([0, 0], [1, 1], [11, 9],
                                               [("periodic", None), ("periodic", None), ("dirichlet", 0), ("dirichlet", 0)], dt)),
            op_case(f"1d_neumann_{tag}", dspec([-pi / 2], [pi / 4], [13], [("neumann", -0.25), ("dirichlet", -0.5)], dt)),
            op_case(f"1d_periodic_{tag}", dspec([0], [1], [12], [("periodic", None), ("periodic", None)], dt)),
        ]
        print(f"  {len(ops)} cases after {dt}")
    # reset the global default dtype the reference flips (backend.py:31,38)
    torch.set_default_dtype(torch.float64)
    torch.save(ops, os.path.join(HERE, "ops.pt"))

    print("edge fixtures")
    edges = []
    for dt in ("double", "single"):
        tag = "f64" if dt == "double" else "f32"
        edges += [
            edge_case(f"edge_3d_dirichlet_{tag}", dspec([0, 0, 0], [1, 1, 2], [6, 5, 7], [("dirichlet", 0.3)] * 6, dt)),
            edge_case(f"edge_3d_mixed_{tag}", dspec([0, 0, 0], [1, 2, 1.5], [7, 6, 8], mixed3, dt)),
            edge_case(f"edge_2d_{tag}", dspec([0, 0], [1, 0.5], [9, 8],
                                            [("neumann", 0.0), ("dirichlet", 0.0), ("dirichlet", 1.0), ("dirichlet", 1.0)], dt)),
            edge_case(f"edge_1d_{tag}", dspec([0], [1], [12], [("dirichlet", 0.0), ("neumann", 0.5)], dt)),
            edge_case(f"edge_1d_periodic_{tag}", dspec([0], [2], [9], [("periodic", None), ("periodic", None)], dt)),
        ]
    torch.set_default_dtype(torch.float64)
    torch.save(edges, os.path.join(HERE, "edges.pt"))
    print(f"  {len(edges)} edge cases")

    print("rz fixtures")
    rzs = []
    for dt in ("double", "single"):
        tag = "f64" if dt == "double" else "f32"
        rz_a = dict(dspec([0, 0], [1, 1], [9, 11], [("neumann", 0.0), ("dirichlet", "rz_bc_ru"), ("dirichlet", "rz_bc_zl"),
                                                   ("dirichlet", "rz_bc_zu")], dt), rz=True)
        rz_b = dict(dspec([0.5, 0], [2, 1], [10, 8], [("dirichlet", 0.5), ("neumann", 0.3), ("symmetry", None),
                                                     ("neumann", -0.2)], dt), rz=True)
        rz_c = dict(dspec([0, -1], [1, 1], [8, 9], [("dirichlet", 0.0)] * 4, dt), rz=True)
        rzs += [rz_op_case(f"rz_axis_{tag}", rz_a), rz_op_case(f"rz_offaxis_{tag}", rz_b), rz_op_case(f"rz_dirichlet_{tag}", rz_c)]
    torch.set_default_dtype(torch.float64)
    torch.save(rzs, os.path.join(HERE, "rz_ops.pt"))
    print(f"  {len(rzs)} rz cases")

    print("solver fixtures")
    L1 = [("laplacian", 1.0, 1.0)]
    sol = []
    # config 1 of BASELINE.json: 64x64 Dirichlet via poisson_2d_bc
    sol.append(solver_case("cfg1_2d_64_cg", dspec([0, 0], [1, 1], [64, 64], [("dirichlet", "poisson_2d_bc")] * 4),
                           L1, "poisson_nd", "cg", 1e-6, 1000))
    sol.append(solver_case("cfg1_2d_64_bicgstab", dspec([0, 0], [1, 1], [64, 64], [("dirichlet", "poisson_2d_bc")] * 4),
                           L1, "poisson_nd", "bicgstab", 1e-6, 1000))
    sol.append(solver_case("nb_2d_100_cg", dspec([0, 0], [1, 1], [100, 100], [("dirichlet", "poisson_2d_bc")] * 4),
                           L1, "poisson_nd", "cg", 1e-6, 1000))
    # tests/test_solver.py:30-88
    for m in ("cg", "bicgstab"):
        sol.append(solver_case(f"t_1d_11_{m}", dspec([0], [1], [11], [("dirichlet", "poisson_1d_bc")] * 2),
                               L1, "poisson_nd", m, 1e-6, 1000))
        sol.append(solver_case(f"t_2d_101_{m}", dspec([0, 0], [1, 1], [101, 101], [("dirichlet", "poisson_2d_bc")] * 4),
                               L1, "poisson_nd", m, 1e-6, 1000))
        sol.append(solver_case(f"t_3d_11_{m}", dspec([0, 0, 0], [1, 1, 1], [11, 11, 11], [D0] * 6),
                               L1, "poisson_nd", m, 1e-6, 1000))
    # tests/test_solver.py:91-151 heat conduction, laplacian(var) == 0.0
    sol.append(solver_case("t_heat_2d_11_bicgstab", dspec([0, 0], [1, 1], [11, 11],
                           [("neumann", 0.0), ("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 1.0)]),
                           [("laplacian", 1.0, None)], 0.0, "bicgstab", 1e-8, 1000))

    # tests/test_solver.py:164-207  -laplacian(var) == rhs, periodic x
    def rhs_periodic(mesh, var):
        r = torch.zeros_like(var())
        r[0] = mesh.X * torch.sin(5.0 * pi * mesh.Y) + torch.exp(-((mesh.X - 0.5) ** 2 + (mesh.Y - 0.5) ** 2) / 0.02)
        return r

    sol.append(solver_case("t_periodic_2d_101_bicgstab", dspec([0, 0], [1, 1], [101, 101],
                           [("periodic", None), ("periodic", None), ("dirichlet", 0), ("dirichlet", 0)]),
                           [("laplacian", -1.0, None)], rhs_periodic, "bicgstab", 1e-8, 1000))

    # tests/test_solver.py:210-268
    def rhs_1dn(mesh, var):
        r = torch.zeros_like(var())
        r[0] = torch.cos(pi / 2 * mesh.X + pi / 4)
        return r

    sol.append(solver_case("t_neumann_1d_101_bicgstab", dspec([-pi / 2], [pi / 4], [101],
                           [("neumann", -1 / 4), ("dirichlet", -1 / 2)]),
                           L1, rhs_1dn, "bicgstab", 1e-6, 1000))

    # tests/test_solver.py:271-306
    def rhs_2dn(mesh, var):
        r = torch.zeros_like(var())
        r[0] = -2 * pi**2 * torch.sin(pi * mesh.X) * torch.sin(pi * mesh.Y)
        return r

    sol.append(solver_case("t_neumann_2d_101_cg", dspec([0, 0], [0.5, 0.5], [101, 101],
                           [("dirichlet", 0), ("neumann", 0), ("dirichlet", 0), ("neumann", 0)]),
                           L1, rhs_2dn, "cg", 1e-6, 1000))
    # demos/advection_diffusion notebook + tests/test_solver.py:361-390: grad(var) - laplacian(eps, var) == 1.0
    for eps in (1.0, 0.5, 0.2, 0.1, 0.02):
        sol.append(solver_case(f"nb_advdiff_1d_51_eps{eps}", dspec([0], [1], [51], [D0] * 2),
                               [("grad", 1.0, None), ("laplacian", -1.0, eps)], 1.0, "bicgstab", 1e-5, 1000, init=0.5))
    # 3-D random-rhs cases
    for n, m in ((32, "cg"), (32, "bicgstab"), (64, "cg")):
        sol.append(solver_case(f"rand_3d_{n}_{m}", dspec([0, 0, 0], [1, 1, 1], [n, n, n], [D0] * 6),
                               L1, ("rand", 1234), m, 1e-8, 5000, keep_solution=(n <= 32)))
    sol.append(solver_case("rand_3d_32_mixed_bicgstab", dspec([0, 0, 0], [1, 1, 1], [32, 32, 32], mixed3),
                           L1, ("rand", 1234), "bicgstab", 1e-8, 5000))
    # non-converging CG on the non-symmetric mixed operator: max_it+1 iterations (linalg.py:144-150)
    sol.append(solver_case("rand_3d_16_mixed_cg_maxit", dspec([0, 0, 0], [1, 1, 1], [16, 16, 16], mixed3),
                           L1, ("rand", 1234), "cg", 1e-8, 25))
    sol.append(solver_case("rand_3d_16_bicgstab_maxit", dspec([0, 0, 0], [1, 1, 1], [16, 16, 16], [D0] * 6),
                           L1, ("rand", 1234), "bicgstab", 1e-30, 10))
    for m_it in (5, 20):
        sol.append(solver_case(f"rand_3d_24_mixed_bicgstab_it{m_it}", dspec([0, 0, 0], [1, 1, 1], [24, 20, 28], mixed3),
                               L1, ("rand", 1234), "bicgstab", 1e-30, m_it))
        sol.append(solver_case(f"rand_2d_40_bicgstab_it{m_it}", dspec([0, 0], [1, 2], [40, 36],
                               [("neumann", 0.3), ("dirichlet", 1.0), ("dirichlet", 0.0), ("symmetry", None)]),
                               L1, ("rand", 77), "bicgstab", 1e-30, m_it))
    # upwind-div + laplacian steady problem through BiCGSTAB
    sol.append(solver_case("advdiff_2d_33_bicgstab", dspec([0, 0], [1, 1], [33, 33], [D0] * 4),
                           [("div", 1.0, 0.5), ("laplacian", -1.0, 0.1)], ("rand", 1234), "bicgstab", 1e-8, 2000,
                           div_cfg={"div": {"limiter": "upwind", "edge": False}}))
    def rhs_rz(mesh, var):
        r = torch.zeros_like(var())
        r[0] = -torch.sin(mesh.X) / (mesh.X * torch.exp(mesh.Z))
        r[0][mesh.X.eq(0.0)] = -1.0 / torch.exp(mesh.Z[mesh.X.eq(0.0)])
        return r

    rz_bcs = [("neumann", 0.0), ("dirichlet", "rz_bc_ru"), ("dirichlet", "rz_bc_zl"), ("dirichlet", "rz_bc_zu")]
    sol.append(solver_case("t_poisson_rz_101_bicgstab", dict(dspec([0, 0], [1, 1], [101, 101], rz_bcs), rz=True),
                           L1, rhs_rz, "bicgstab", 1e-5, 1000))
    sol.append(solver_case("rz_33_bicgstab_it12", dict(dspec([0, 0], [1, 1], [33, 29], rz_bcs), rz=True),
                           L1, rhs_rz, "bicgstab", 1e-30, 12))
    # fp32: global default dtype flips to float32 inside the reference
    sol.append(solver_case("cfg1_2d_64_cg_f32", dspec([0, 0], [1, 1], [64, 64], [("dirichlet", "poisson_2d_bc")] * 4, "single"),
                           L1, "poisson_nd", "cg", 1e-6, 1000))
    sol.append(solver_case("rand_2d_32_bicgstab_f32", dspec([0, 0], [1, 1], [32, 32], [D0] * 4, "single"),
                           L1, ("rand", 1234), "bicgstab", 1e-4, 1000))
    torch.set_default_dtype(torch.float64)
    torch.save(sol, os.path.join(HERE, "solvers.pt"))
    for f in ("ops.pt", "edges.pt", "rz_ops.pt", "solvers.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) / 1e6, "MB")


if __name__ == "__main__":
    main()
