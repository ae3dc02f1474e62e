"""Shared-memory-resident kernels for 2-D grids (csrc/kernels_resident.cuh): explicit Euler stepping and the
whole-solve CG as one cooperative launch.  Euler must be bit-identical to the oracle (small shapes) and to the
streaming star engine (BASELINE config 3 size 1024^2); CG must match the fused kernels and the oracle in iteration
count, tolerance (1e-10) and solution (1e-9 relative).  Shapes cover one row per CTA, a short last CTA, rows that
are not a multiple of the thread walk, and more CTAs than rows."""
from __future__ import annotations

import os
import warnings

import pytest
import torch

from oracle import fd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _euler_variant:
    """"resident" (default: uniform-coefficient fast path where it applies), "items" (resident, general item loop),
    "stream" (per-step launches of the star engine)."""

    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.prev = {k: os.environ.get(k) for k in ("PA_EULER_VARIANT", "PA_RES_PATH")}
        os.environ.pop("PA_EULER_VARIANT", None)
        os.environ.pop("PA_RES_PATH", None)
        if self.v == "stream":
            os.environ["PA_EULER_VARIANT"] = "stream"
        elif self.v == "items":
            os.environ["PA_RES_PATH"] = "items"

    def __exit__(self, *a):
        for k, v in self.prev.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _euler_run(shape, limiter, n_steps, dtype="double", with_rhs=False, seed=1234):
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, DEV, dtype)
    vals = [0.0, 0.25, -0.5, 1.0][: 2 * nd]
    var = Field("c", 1, mesh, {"domain": mixed_bcs(vals, ["dirichlet"] * (2 * nd)), "obstacle": None})
    g = torch.Generator().manual_seed(seed)
    phi0 = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(var().dtype)
    src = (torch.rand(1, *shape, generator=g, dtype=torch.float64) - 0.5).to(var().dtype) if with_rhs else None
    var.set_var_tensor(phi0.to(DEV))
    nu, u = 0.1, 1.0
    dt = 0.2 * min(mesh._dx) ** 2 / nu
    var.set_time(dt, 0.0)
    fdm = FDM({"div": {"limiter": limiter, "edge": False}})
    solver = Solver({"fdm": {"method": "euler", "tol": 0.0, "max_it": 0, "report": False, "n_steps": n_steps}})
    solver.set_eq(fdm.ddt(var) + fdm.div(u, var) - fdm.laplacian(nu, var) == (src.to(DEV) if with_rhs else 0.0))
    solver.solve()
    torch.set_default_dtype(torch.float64)
    return mesh, var, phi0, src, dt, vals


@pytest.mark.parametrize("limiter", ["upwind", "upwind_fd"])
@pytest.mark.parametrize("shape,n_steps", [([40, 64], 7), ([7, 16], 4), ([150, 128], 5), ([149, 32], 2),
                                           ([300, 96], 6), ([3, 8], 3), ([449, 1032], 3)])
@pytest.mark.parametrize("with_rhs", [False, True])
@pytest.mark.parametrize("path", ["resident", "items"])
def test_euler_resident_vs_oracle(limiter, shape, n_steps, with_rhs, path):
    with _euler_variant(path):
        mesh, var, phi0, src, dt, vals = _euler_run(shape, limiter, n_steps, with_rhs=with_rhs)
    nd = len(shape)
    xs, dx = O.make_axes([0.0] * nd, [1.0] * nd, shape)
    bcs = [O.FaceBC(f, "dirichlet", v) for f, v in zip(O.FACES[: 2 * nd], vals)]
    x = phi0.clone()
    eq = O.Equation([O.Term("div", 1.0, 1.0, limiter), O.Term("laplacian", -1.0, 0.1)], dx, xs, bcs).build(x)
    prev = x
    for _ in range(n_steps):
        prev = x
        x = O.euler_step(eq, x, src, dt)
    assert torch.equal(var().cpu(), x), (var().cpu() - x).abs().max().item()
    # the loser of the ping-pong is the step before (the reference's VARo)
    assert torch.equal(var.VARo.cpu(), prev), (var.VARo.cpu() - prev).abs().max().item()


@pytest.mark.parametrize("dtype", ["double", "single"])
@pytest.mark.parametrize("shape,n_steps", [([1024, 1024], 41), ([512, 2048], 12), ([1000, 520], 9), ([512, 512], 30),
                                           ([256, 256], 30), ([300, 128], 11)])
def test_euler_resident_equals_stream(dtype, shape, n_steps):
    """BASELINE config 3 size: the resident launch and the per-step star-engine launches agree bit for bit."""
    out = {}
    for v in ("resident", "items", "stream"):
        with _euler_variant(v):
            _, var, *_ = _euler_run(shape, "upwind", n_steps, dtype=dtype)
        out[v] = (var().clone(), var.VARo.clone())
    assert torch.isfinite(out["stream"][0]).all()
    for v in ("resident", "items"):
        assert torch.equal(out[v][0], out["stream"][0]), v
        assert torch.equal(out[v][1], out["stream"][1]), v


def _cg_run(shape, variant, max_it, tol=1e-30, dtype="double", seed=1234, vals=None):
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, DEV, dtype)
    vals = vals or [0.0, 0.5, -1.0, 2.0][: 2 * nd]
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, ["dirichlet"] * (2 * nd)), "obstacle": None})
    g = torch.Generator().manual_seed(seed)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(var().dtype).to(DEV)
    s = Solver({"fdm": {"method": "cg", "tol": tol, "max_it": max_it, "report": False, "variant": variant}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = s.solve()
    torch.set_default_dtype(torch.float64)
    return var, rep, vals


@pytest.mark.parametrize("shape,max_it", [([64, 64], 30), ([150, 128], 25), ([149, 32], 25), ([7, 16], 10),
                                          ([300, 96], 40), ([1024, 1024], 60), ([449, 1032], 15)])
@pytest.mark.parametrize("path", ["resident", "items"])
def test_cg_resident_equals_fused(shape, max_it, path):
    with _euler_variant(path):
        a, rep_a, _ = _cg_run(shape, 6, max_it)
    b, rep_b, _ = _cg_run(shape, 4, max_it)
    assert rep_a["itr"] == rep_b["itr"] == max_it + 1, (rep_a, rep_b)
    assert abs(rep_a["tol"] - rep_b["tol"]) <= 1e-10 * max(1.0, abs(rep_b["tol"])), (rep_a, rep_b)
    smax = b().abs().max().item()
    assert (a() - b()).abs().max().item() <= 1e-9 * smax
    assert (a.VARo - b.VARo).abs().max().item() <= 1e-9 * smax
    assert a._last_launches < b._last_launches  # one launch instead of two per iteration


@pytest.mark.parametrize("shape", [[40, 64], [96, 48], [130, 256]])
@pytest.mark.parametrize("path", ["resident", "items"])
def test_cg_resident_converged_vs_oracle(shape, path):
    with _euler_variant(path):
        var, rep, vals = _cg_run(shape, 6, 2000, tol=1e-8)
    nd = len(shape)
    xs, dx = O.make_axes([0.0] * nd, [1.0] * nd, shape)
    bcs = [O.FaceBC(f, "dirichlet", v) for f, v in zip(O.FACES[: 2 * nd], vals)]
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64)
    x0 = torch.zeros(1, *shape, dtype=torch.float64)
    eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(x0)
    sol, rep_o, x_prev = O.cg(eq, x0, eq.adjust_rhs(x0, rhs.clone()), 1e-8, 2000)
    assert rep["itr"] == rep_o["itr"], (rep, rep_o)
    assert abs(rep["tol"] - rep_o["tol"]) <= 1e-10
    scale = sol.abs().max().item()
    assert (var().cpu() - sol).abs().max().item() <= 1e-9 * scale
    assert (var.VARo.cpu() - x_prev).abs().max().item() <= 1e-9 * scale


def test_cg_resident_fp32_and_auto():
    """fp32 instantiation, and `auto` takes the resident kernel on a 2-D grid above the tiny-grid limit."""
    a, rep_a, _ = _cg_run([512, 512], 6, 30, dtype="single")
    b, rep_b, _ = _cg_run([512, 512], 4, 30, dtype="single")
    assert rep_a["itr"] == rep_b["itr"]
    assert (a() - b()).abs().max().item() <= 1e-4 * b().abs().max().item()
    c, rep_c, _ = _cg_run([512, 512], 0, 30)
    d, rep_d, _ = _cg_run([512, 512], 6, 30)
    assert c._last_launches == d._last_launches
    assert torch.equal(c(), d())


def _jacobi_run(shape, variant, max_it, tol=1e-300, dtype="double", seed=77):
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, DEV, dtype)
    vals = [0.0, 0.5, -1.0, 2.0][: 2 * nd]
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, ["dirichlet"] * (2 * nd)), "obstacle": None})
    g = torch.Generator().manual_seed(seed)
    rhs = (torch.rand(1, *shape, generator=g, dtype=torch.float64) - 0.5).to(var().dtype)
    s = Solver({"fdm": {"method": "jacobi", "tol": tol, "max_it": max_it, "report": False, "variant": variant}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs.to(DEV))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = s.solve()
    torch.set_default_dtype(torch.float64)
    return var, rep, vals, rhs


@pytest.mark.parametrize("shape,sweeps", [([64, 64], 9), ([150, 128], 6), ([300, 1024], 5), ([1024, 1024], 12),
                                          ([7, 16], 4), ([449, 512], 7)])
def test_jacobi_resident_vs_oracle_and_stream(shape, sweeps):
    """The resident Jacobi solve (one cooperative launch) against the oracle -- bit-identical iterate after a fixed
    number of sweeps, same sweep count -- and against the streaming sweeps (variant 4) incl. the previous iterate."""
    a, rep_a, vals, rhs = _jacobi_run(shape, 6, sweeps - 1)
    b, rep_b, _, _ = _jacobi_run(shape, 4, sweeps - 1)
    assert rep_a["itr"] == rep_b["itr"] == sweeps, (rep_a, rep_b)
    assert a._last_launches < b._last_launches  # one launch instead of one per sweep
    assert torch.equal(a(), b())
    assert torch.equal(a.VARo, b.VARo)
    assert abs(rep_a["tol"] - rep_b["tol"]) <= 1e-12 * max(1.0, abs(rep_b["tol"]))
    if shape[0] * shape[1] <= 200000:
        nd = len(shape)
        xs, dx = O.make_axes([0.0] * nd, [1.0] * nd, shape)
        bcs = [O.FaceBC(f, "dirichlet", v) for f, v in zip(O.FACES[: 2 * nd], vals)]
        x0 = torch.zeros(1, *shape, dtype=torch.float64)
        eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(x0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol, rep_o, _ = O.jacobi(eq, x0, eq.adjust_rhs(x0, rhs.clone()), 1e-300, sweeps - 1)
        assert rep_o["itr"] == sweeps
        assert torch.equal(a().cpu(), sol), (a().cpu() - sol).abs().max().item()


def test_jacobi_resident_converges_like_stream():
    a, rep_a, _, _ = _jacobi_run([96, 64], 6, 20000, tol=1e-6)
    b, rep_b, _, _ = _jacobi_run([96, 64], 4, 20000, tol=1e-6)
    assert rep_a["converge"] and rep_a["itr"] == rep_b["itr"], (rep_a, rep_b)
    assert torch.equal(a(), b())


@pytest.mark.parametrize("method", ["cg", "jacobi"])
def test_implicit_euler_through_resident_kernels(method):
    """fdm.ddt + a linear-solver method = implicit Euler: (1/dt) phi' - nu lap(phi') = rhs + (1/dt) phi.  On a 2-D
    Dirichlet grid the resident CG takes the 1/dt term as its shift (OpDev::shift), the resident Jacobi as a second
    operator; both must reproduce the streaming kernels (variant 4) step for step."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    shape = [300, 512]
    out = {}
    for variant in (6, 4):
        mesh = Mesh(Box([0.0, 0.0], [1.0, 1.0]), None, shape, DEV, "double")
        var = Field("c", 1, mesh, {"domain": mixed_bcs([0.0, 1.0, 0.5, -0.25], ["dirichlet"] * 4), "obstacle": None})
        g = torch.Generator().manual_seed(99)
        var.set_var_tensor(torch.rand(1, *shape, generator=g, dtype=torch.float64).to(DEV))
        src = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(DEV)
        nu = 0.1
        var.set_time(5.0 * min(mesh._dx) ** 2 / nu, 0.0)
        tol, max_it = (1e-9, 4000) if method == "cg" else (1e-300, 30)
        solver = Solver({"fdm": {"method": method, "tol": tol, "max_it": max_it, "report": False, "n_steps": 2,
                                 "variant": variant}})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fdm = FDM()
            solver.set_eq(fdm.ddt(var) - fdm.laplacian(nu, var) == src)
            rep = solver.solve()
        out[variant] = (rep, var().clone(), var._last_launches)
        torch.set_default_dtype(torch.float64)
    (ra, xa, la), (rb, xb, lb) = out[6], out[4]
    assert ra["itr"] == rb["itr"], (ra, rb)
    assert la < lb  # whole-solve launches
    if method == "jacobi":
        assert torch.equal(xa, xb)
    else:
        assert abs(ra["tol"] - rb["tol"]) <= 1e-10
        assert (xa - xb).abs().max().item() <= 1e-9 * xb.abs().max().item()
