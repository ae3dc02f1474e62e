"""Shared helpers for the parity tests: load golden fixtures, build oracle objects."""
from __future__ import annotations

import os

import torch

from oracle import fd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TDTYPE = {"double": torch.float64, "single": torch.float32}


def load(name: str):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def oracle_bcs(case) -> list[O.FaceBC]:
    return [O.FaceBC(f, k, v, rz=bool(case["spec"].get("rz"))) for f, k, v in case["bcs"]]


def oracle_axes(case):
    spec = case["spec"]
    xs, dx = O.make_axes(spec["lower"], spec["upper"], spec["nx"], TDTYPE[spec["dtype"]])
    assert dx == case["dx"], (dx, case["dx"])
    return xs, dx


def case_rhs(case, shape, dtype):
    rhs = case["rhs"]
    if isinstance(rhs, tuple) and rhs[0] == "rand":
        g = torch.Generator().manual_seed(rhs[1])
        return torch.rand(shape, generator=g, dtype=torch.float64).to(dtype)
    if isinstance(rhs, torch.Tensor):
        return rhs.clone()
    return torch.zeros(shape, dtype=dtype) + float(rhs)


def is_rz(case) -> bool:
    return bool(case["spec"].get("rz"))


def oracle_terms(case):
    limiter = (case.get("div_cfg") or {}).get("div", {}).get("limiter", "none")
    return [O.Term(kind, sign=float(sign), param=param, limiter=limiter if kind == "div" else "none")
            for kind, sign, param in case["terms"]]


# ---- product-side builders (pyapes_b200 public API) ----------------------------------------
def product_field(case, device="cuda", init=0.0):
    """Mesh + Field of the product package from a fixture's spec / frozen BC list."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.variables import Field

    spec = case["spec"]
    if spec.get("rz"):
        from pyapes_b200.geometry import Cylinder

        geo = Cylinder(list(spec["lower"]), list(spec["upper"]))
    else:
        geo = Box(list(spec["lower"]), list(spec["upper"]))
    mesh = Mesh(geo, None, list(spec["nx"]), device, spec["dtype"])
    cfg = []
    for face, kind, val in case["bcs"]:
        if isinstance(val, torch.Tensor):
            val = val.to(device)
        cfg.append({"bc_face": face, "bc_type": kind, "bc_val": val, "bc_val_opt": None})
    var = Field("p", 1, mesh, {"domain": cfg, "obstacle": None})
    if init != 0.0:
        var.set_var_tensor(torch.zeros_like(var()) + init)
    return mesh, var


def product_equation(case, fdm, var, dev):
    eq = None
    for kind, sign, param in case["terms"]:
        if isinstance(param, torch.Tensor):
            param = param.to(dev)
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
    return eq


# ---- the explicit operators of the tile fixtures (tests/golden/make_golden_tiles.py) ---------
def oracle_tile_outputs(case, phi=None):
    """Every output key of an ops_tiles.pt case, computed by the oracle."""
    xs, dx = oracle_axes(case)
    bcs = oracle_bcs(case)
    phi = case["phi"].clone() if phi is None else phi
    u = case["u_const"]
    has_ns = any(k in ("neumann", "symmetry") for _, k, _ in case["bcs"])
    out = {}
    out["lap"] = O.Equation([O.Term("laplacian", 1.0, None)], dx, xs, bcs).build(phi).aop(phi)
    grad = O.apply_grad(O.grad_coeffs(phi, dx, bcs), phi)
    out["grad"] = grad
    out["div_upwind_const"] = O.apply_scalar_op(O.div_coeffs(u, phi, dx, bcs, "upwind"), phi)
    if not has_ns:
        out["div_central_const"] = O.apply_scalar_op(O.div_coeffs(u, phi, dx, bcs, "none"), phi)
    e = O.Equation([O.Term("div", 1.0, u, "upwind"), O.Term("laplacian", -1.0, 0.1)], dx, xs, bcs).build(phi)
    out["advdiff"] = e.aop(phi)
    out["lap_edge"] = O.edge_laplacian(O.apply_scalar_op(O.laplacian_coeffs(phi, dx, bcs), phi), phi, dx)
    out["grad_edge"] = O.edge_grad(grad.clone(), phi, dx)
    return out


def product_tile_outputs(case, var):
    """The same keys through the public API of the product package (-> C ABI -> CUDA)."""
    import torch as _t

    from pyapes_b200.solver.fdc import FDC
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver

    u = case["u_const"]
    has_ns = any(k in ("neumann", "symmetry") for _, k, _ in case["bcs"])
    out = {}
    s = Solver(None)
    s.set_eq(FDM().laplacian(var) == _t.zeros_like(var()))
    out["lap"] = s.Aop(var)
    out["grad"] = FDC({"grad": {"edge": False}}).grad(var)
    out["div_upwind_const"] = FDC({"div": {"limiter": "upwind", "edge": False}}).div(u, var)
    if not has_ns:
        out["div_central_const"] = FDC({"div": {"limiter": "none", "edge": False}}).div(u, var)
    fdm = FDM({"div": {"limiter": "upwind", "edge": False}})
    s = Solver(None)
    s.set_eq(fdm.div(u, var) - fdm.laplacian(0.1, var) == _t.zeros_like(var()))
    out["advdiff"] = s.Aop(var)
    out["lap_edge"] = FDC({"laplacian": {"edge": True}}).laplacian(var)
    out["grad_edge"] = FDC({"grad": {"edge": True}}).grad(var)
    FDC({"laplacian": {"edge": False}, "grad": {"edge": False}, "div": {"limiter": "none", "edge": False}})
    return out
