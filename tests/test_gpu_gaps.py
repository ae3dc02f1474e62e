"""Parity branches the first round left untested ("close the parity gaps"):
 (a) BASELINE config 1 exactly as a user writes it -- `poisson_bcs(2)` CALLABLES through Field -> itr 134,
     tol 8.200499904126584e-07 -- plus callable Neumann / Dirichlet fixtures from the real
     reference, and the var-dependent callable that must be refused;
 (c) the NaN/Inf tolerance -> RuntimeError path (linalg.py:334-336), with the reference's post-mortem state;
 (d) pa_cg_solve_host straight through ctypes with numpy host buffers;
 (e) the reference's own tests/test_solver.py, unmodified, against this package (install_as_pyapes).
((b), fp32 tol/solution, lives in test_gpu_parity.py::test_solver_fixtures.)"""
import ctypes as C
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

from oracle import fd_oracle as O
from tests import _util as U
from tests.golden.bc_callables import CALLABLES

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALL = U.load("callables.pt")
JACDIV = U.load("jacdiv.pt")
SOL = {c["name"]: c for c in U.load("solvers.pt")}


def test_config1_as_written_callable_dirichlet():
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.testing.poisson import poisson_bcs, poisson_exact_nd, poisson_rhs_nd
    from pyapes_b200.variables import Field

    mesh = Mesh(Box[0:1, 0:1], None, [64, 64], DEV, "double")
    var = Field("p", 1, mesh, {"domain": poisson_bcs(2), "obstacle": None})
    assert all(callable(bc.bc_val) for bc in var.bcs)
    rhs = poisson_rhs_nd(mesh, var)
    solver = Solver({"fdm": {"method": "cg", "tol": 1e-6, "max_it": 1000, "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == rhs)
    rep = solver.solve()
    assert rep["itr"] == 134 and rep["converge"], rep
    assert abs(rep["tol"] - 8.200499904126584e-07) <= 1e-10, rep
    ref = SOL["cfg1_2d_64_cg"]
    assert rep["itr"] == ref["report"]["itr"]
    sol = var().cpu()
    assert (sol - ref["solution"]).abs().max().item() <= 1e-9 * ref["solution"].abs().max().item()
    assert (sol[0] - poisson_exact_nd(mesh).cpu()).abs().max().item() < 2e-7  # 1.74e-07


def _callable_field(case):
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    spec = case["spec"]
    mesh = Mesh(Box(list(spec["lower"]), list(spec["upper"])), None, list(spec["nx"]), DEV, spec["dtype"])
    vals = [CALLABLES[v] if isinstance(v, str) else v for _, v in spec["bcs"]]
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, [k for k, _ in spec["bcs"]]), "obstacle": None})
    return mesh, var


@pytest.mark.parametrize("case", CALL, ids=[c["name"] for c in CALL])
def test_callable_bc_fixtures(case):
    """Neumann and Dirichlet faces whose value is a callable of (grid, mask) (bcs.py:203-205,240-241)."""
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.linalg import _apply_bc_otf
    from pyapes_b200.solver.ops import Solver

    mesh, var = _callable_field(case)
    out = case["out"]
    var.set_var_tensor(case["phi"].to(DEV).clone())
    s = Solver(None)
    r0 = torch.zeros_like(var())
    s.set_eq(FDM().laplacian(1.0, var) == r0)
    assert torch.equal(s.Aop(var).cpu(), out["lap"])
    assert torch.equal(r0.cpu(), out["lap_rhs_adj"])
    _apply_bc_otf(var, mesh)
    assert torch.equal(var().cpu(), out["bc_applied"])

    var.set_var_tensor(torch.zeros_like(var()))
    solver = Solver({"fdm": {"method": case["method"], "tol": case["tol"], "max_it": case["max_it"], "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == case["rhs"].to(DEV).clone())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = solver.solve()
    ref, sol = case["report"], var().cpu()
    smax = case["solution"].abs().max().item()
    if case["method"] == "cg" or ref["itr"] >= case["max_it"]:  # CG, or BiCGSTAB in lockstep
        assert rep["itr"] == ref["itr"], (rep, ref)
        assert abs(rep["tol"] - ref["tol"]) <= 1e-7 * ref["tol"] + 1e-10, (rep, ref)
        assert (sol - case["solution"]).abs().max().item() <= 1e-9 * smax
    else:  # converged BiCGSTAB: the reference's own count moves by ~10 % under 1-ulp noise
        assert rep["converge"] and abs(rep["itr"] - ref["itr"]) <= max(3, ref["itr"] // 8), (rep, ref)
        assert (sol - case["solution"]).abs().max().item() <= 1e-6 * smax


def test_var_dependent_callable_is_refused_in_solvers_but_applies_directly():
    """The reference re-evaluates a callable on every iterate; the device-resident iteration cannot, so a
    callable that reads `var` raises instead of silently freezing; BC.apply itself, which
    evaluates at every call like the reference, still takes it."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    mesh = Mesh(Box[0:1, 0:1], None, [20, 24], DEV, "double")
    vals = [CALLABLES["dirichlet_of_var"], 0.0, 0.0, 0.0]
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, ["dirichlet"] * 4), "obstacle": None})
    g = torch.Generator().manual_seed(2)
    phi = torch.rand(1, 20, 24, generator=g, dtype=torch.float64).to(DEV)
    var.set_var_tensor(phi.clone())
    x = phi.clone()
    var.bcs[0].apply(x, mesh.grid, 0)
    assert torch.equal(x[0, 0], 0.5 * phi[0, 0] + 1.0) and torch.equal(x[0, 1:], phi[0, 1:])
    solver = Solver({"fdm": {"method": "cg", "tol": 1e-6, "max_it": 10, "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == torch.zeros_like(var()))
    with pytest.raises(NotImplementedError, match="depends on the field"):
        solver.solve()


@pytest.mark.parametrize("method", ["cg", "bicgstab"])
@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_invalid_tolerance_raises_like_the_reference(method, bad):
    """linalg.py:334-336.  Probed on the real reference (16^2 Dirichlet, one bad RHS entry): both solvers
    raise RuntimeError('Invalid tolerance detected!'); CG has already written its first update (the field
    holds NaN), BiCGSTAB fails before any update (the field is still the initial zeros)."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    for variant in (0, 1):
        mesh = Mesh(Box[0:1, 0:1], None, [16, 16], DEV, "double")
        var = Field("p", 1, mesh, {"domain": homogeneous_bcs(2, 0.0, "dirichlet"), "obstacle": None})
        g = torch.Generator().manual_seed(3)
        rhs = torch.rand(1, 16, 16, generator=g, dtype=torch.float64)
        rhs[0, 5, 5] = bad
        s = Solver({"fdm": {"method": method, "tol": 1e-6, "max_it": 50, "report": False, "variant": variant}})
        s.set_eq(FDM().laplacian(1.0, var) == rhs.to(DEV))
        with pytest.raises(RuntimeError, match="Invalid tolerance detected"):
            s.solve()
        finite = bool(torch.isfinite(var()).all())
        if method == "cg":
            assert not finite
        else:
            assert finite and var().abs().sum().item() == 0.0


def test_fp32_bicgstab_stagnation_case_of_the_survey():
    """fp32 BiCGSTAB on config 1 stagnates in the reference and ends in 'Invalid tolerance
    detected' after ~425 iterations, the field still finite.  Here the dot products are accumulated in fp64,
    so the run may instead converge or stop at max_it; whichever way it ends, it must end like the reference
    API does: converge / RuntimeWarning / that RuntimeError -- and leave a finite field."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.testing.poisson import poisson_bcs, poisson_rhs_nd
    from pyapes_b200.variables import Field

    mesh = Mesh(Box[0:1, 0:1], None, [64, 64], DEV, "single")
    var = Field("p", 1, mesh, {"domain": poisson_bcs(2), "obstacle": None})
    solver = Solver({"fdm": {"method": "bicgstab", "tol": 1e-6, "max_it": 1000, "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == poisson_rhs_nd(mesh, var))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            rep = solver.solve()
            ended = "converged" if rep["converge"] else "max_it"
            if ended == "max_it":
                assert any("Maximum iteration" in str(x.message) for x in w)
        except RuntimeError as e:
            assert "Invalid tolerance detected" in str(e)
            ended = "invalid tolerance"
    torch.set_default_dtype(torch.float64)
    assert bool(torch.isfinite(var()).all()), ended
    print("fp32 BiCGSTAB config 1 ended with:", ended)


def test_pa_cg_solve_host_numpy_buffers():
    """The host-buffer plugin entry: numpy arrays in, solution out, no torch tensor on the call."""
    from oracle import fd_oracle as O
    from pyapes_b200 import _lower as L
    from pyapes_b200 import _native as N

    n = [24, 20, 32]
    lib = N.lib()
    g = N.Grid()
    for a in range(3):
        g.n[a], g.lo[a], g.hi[a] = n[a], 1, n[a] - 1
    g.gn0, g.goff0, g.olo0, g.ohi0, g.ndim = n[0], 0, 0, n[0], 3
    dx = [1.0 / (v - 1) for v in n]
    star = L.laplacian_star(n, dx, [], torch.float64)
    op, keep = L.lower_op(star, 3, torch.float64, sign=1.0, param=1.0)
    eq = N.Equation()
    eq.nops = 1
    eq.ops[0] = op
    vals = [0.0, 1.0, 0.5, 0.0, -0.25, 0.0]
    faces = (N.FaceBC * 6)()
    for f in range(6):
        faces[f].axis, faces[f].side, faces[f].kind, faces[f].value = f // 2, (-1 if f % 2 == 0 else 1), 1, vals[f]
    rng = np.random.default_rng(11)
    rhs = rng.random(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.float64)
    cfg = N.SolverCfg()
    cfg.tol, cfg.max_it, cfg.check_every, cfg.use_graph, cfg.variant = 1e-8, 3000, 0, 1, 0
    rep = N.Report()
    N.check(lib.pa_cg_solve_host(g, eq, 6, faces, N.PA_F64, x.ctypes.data_as(C.c_void_p), rhs.ctypes.data_as(C.c_void_p),
                                 cfg, rep))
    xs, odx = O.make_axes([0, 0, 0], [1, 1, 1], n)
    bcs = [O.FaceBC(f, "dirichlet", v) for f, v in zip(O.FACES, vals)]
    x0 = torch.zeros(1, *n, dtype=torch.float64)
    oeq = O.Equation([O.Term("laplacian", 1.0, 1.0)], odx, xs, bcs).build(x0)
    r = oeq.adjust_rhs(x0, torch.from_numpy(rhs.copy())[None])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol, orep, _ = O.cg(oeq, x0, r, 1e-8, 3000)
    assert rep.itr == orep["itr"] and rep.status == N.CONVERGED, (rep.itr, orep)
    assert abs(rep.tol - orep["tol"]) <= 1e-10
    assert np.abs(x - sol[0].numpy()).max() <= 1e-9 * float(sol.abs().max())
    assert rep.result_in_alt == 0
    del keep


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "tests", "test_solver.py")),
                    reason="the reference's test files are not installed under baseline/_ref (build container only)")
def test_reference_own_solver_tests_pass_unmodified():
    """`pytest baseline/_ref/tests/test_solver.py` -- the reference's OWN test file, byte for byte (git-ignored
    copy made by __graft_entry__.install_reference) -- with `pyapes` aliased to this package
    (pyapes_b200.install_as_pyapes via the pytest plugin pyapes_b200.pytest_alias) and the Mesh default
    device redirected to CUDA (PYAPES_B200_DEFAULT_DEVICE)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    env = dict(os.environ, PYAPES_B200_DEFAULT_DEVICE="cuda",
               PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "golden", "_shim")]))
    cmd = [sys.executable, "-m", "pytest", "-p", "pyapes_b200.pytest_alias", "-p", "no:cacheprovider", "-q", "-x",
           "--rootdir", ref, "-c", os.devnull, os.path.join(ref, "tests", "test_solver.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ref, env=env)
    tail = "\n".join((out.stdout + out.stderr).splitlines()[-40:])
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail
    assert "pyapes_alias: pyapes.solver.ops -> pyapes_b200.solver.ops" in out.stdout, tail


@pytest.mark.parametrize("case", JACDIV, ids=[c["name"] for c in JACDIV])
def test_jac_driven_div_fixtures(case):
    """`FDC().div(jac, var)` (fdc.py:639-664,730-735) against the real reference, and the
    reference's own failures for a Hess / for FDM().div(jac, ...)."""
    from pyapes_b200.solver.fdc import FDC, hessian, jacobian
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.variables import Field

    mesh, var = U.product_field(case, DEV)
    var.set_var_tensor(case["phi"].to(DEV).clone())
    other = Field("q", 1, mesh, None)
    other.set_var_tensor(case["q"].to(DEV).clone())
    jac, hess = jacobian(other), hessian(other)
    out = case["out"]
    for lim in ("upwind", "none"):
        if f"div_jac_{lim}" not in out:
            continue
        fdc = FDC({"div": {"limiter": lim, "edge": False}})
        got = fdc.div(jac, var)
        assert torch.equal(got.cpu(), out[f"div_jac_{lim}"]), lim
        assert torch.equal(fdc.div.rhs_adj.cpu(), out[f"div_jac_{lim}_rhs_adj"]), lim
    errors = {"AttributeError": AttributeError, "NotImplementedError": NotImplementedError, "IndexError": IndexError}
    for lim, err in out["hess_errors"].items():
        with pytest.raises(errors[err]):
            FDC({"div": {"limiter": lim, "edge": False}}).div(hess, var)
    assert out["fdm_div_accepts_jac"] is False
    with pytest.raises(AssertionError):
        FDM({"div": {"limiter": "upwind", "edge": False}}).div(jac, var)
    FDC({"laplacian": {"edge": False}, "grad": {"edge": False}, "div": {"limiter": "none", "edge": False}})
    torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("shape", [[48, 40, 64], [40, 1024]])
def test_contraction_mode_is_opt_in_and_within_tolerance(shape):
    """PA_FLAG_CONTRACT (config key "contract"): the fused TMA CG kernels use one FMA where the reference rounds
    twice.  north_star's bar is 1e-12 relative per operator; after ONE iteration (x = alpha d: one operator
    application and two dot products) the iterate must agree with the bit-exact mode to that bar, 20 lockstep
    iterations to 1e-10, and a converged solve must need the same number of iterations +- 1."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    nd = len(shape)
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(DEV)

    def run(tol, max_it, contract):
        mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, DEV, "double")
        var = Field("p", 1, mesh, {"domain": homogeneous_bcs(nd, 0.0, "dirichlet"), "obstacle": None})
        s = Solver({"fdm": {"method": "cg", "tol": tol, "max_it": max_it, "report": False, "variant": 4,
                            "contract": contract}})
        s.set_eq(FDM().laplacian(1.0, var) == rhs.clone())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rep = s.solve()
        return rep, var().clone()

    for max_it, bar in ((0, 1e-12), (19, 1e-10)):
        (re, xe), (rc, xc) = run(1e-30, max_it, False), run(1e-30, max_it, True)
        assert re["itr"] == rc["itr"] == max_it + 1
        scale = xe.abs().max().item()
        assert (xe - xc).abs().max().item() <= bar * scale, (max_it, (xe - xc).abs().max().item() / scale)
        assert abs(re["tol"] - rc["tol"]) <= bar * re["tol"]
        assert not torch.equal(xe, xc) or max_it == 0  # the mode really changes bits (not a no-op flag)
    (re, xe), (rc, xc) = run(1e-8, 5000, False), run(1e-8, 5000, True)
    assert re["converge"] and rc["converge"] and abs(re["itr"] - rc["itr"]) <= 1, (re, rc)
    assert (xe - xc).abs().max().item() <= 1e-8 * xe.abs().max().item()


@pytest.mark.parametrize("shape", [[40, 36, 128], [96, 256]])
@pytest.mark.parametrize("kinds", ["dirichlet", "mixed"])
def test_jacobi_quotient_fallback_is_bit_exact(shape, kinds):
    """The TMA Jacobi sweep divides by a reciprocal plus two FMA corrections when a warp-wide vote finds every residual
    inside a safe exponent window, and falls back to the true division otherwise (kernels_tma_pw.cuh).  A right-hand
    side with patches of exact zeros and of 1e-200 must give the oracle's bits on both routes (lockstep sweeps)."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    nd = len(shape)
    if kinds == "dirichlet":
        ks, vs = ["dirichlet"] * (2 * nd), [0.0, 0.25, -0.5, 1.0, 0.0, 0.5][: 2 * nd]
    elif nd == 3:
        ks, vs = ["dirichlet", "dirichlet", "neumann", "symmetry", "dirichlet", "dirichlet"], [0.0, 0.25, 0.5, None, 0.0, 1.0]
    else:
        ks, vs = ["neumann", "dirichlet", "dirichlet", "dirichlet"], [0.2, 0.0, 1.0, 0.0]
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda", "double")
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vs, ks), "obstacle": None})
    g = torch.Generator().manual_seed(3)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64) - 0.5
    sl = [slice(None)] * (nd + 1)
    sl[1] = slice(2, 9)
    rhs[tuple(sl)] = 0.0          # exact zeros: x = 0 there at first, so the residual is an exact zero
    sl[1] = slice(12, 15)
    rhs[tuple(sl)] *= 1e-200      # below the window (a patch above it would overflow the norm: RuntimeError in both)
    sweeps = 4
    s = Solver({"fdm": {"method": "jacobi", "tol": 1e-300, "max_it": sweeps - 1, "report": False}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs.to("cuda"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = s.solve()
    xs, dx = O.make_axes([0.0] * nd, [1.0] * nd, shape)
    bcs = [O.FaceBC(f, k, v) for f, k, v in zip(O.FACES[: 2 * nd], ks, vs)]
    x0 = torch.zeros(1, *shape, dtype=torch.float64)
    eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(x0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol, rep_o, _ = O.jacobi(eq, x0, eq.adjust_rhs(x0, rhs.clone()), 1e-300, sweeps - 1)
    assert rep["itr"] == rep_o["itr"] == sweeps
    assert torch.equal(var().cpu(), sol), (var().cpu() - sol).abs().max().item()
