"""Measurement helpers (not part of the reference API): per-kernel CUDA-event timings of the CG
iteration, used by bench.py for the roofline object."""
from __future__ import annotations

import ctypes as C

import torch

from pyapes_b200 import _lower as L
from pyapes_b200 import _native as N


def cg_kernel_times(n, iters: int = 20, variant: int = 0, dtype: str = "double", device: str = "cuda") -> dict:
    """n^3 Dirichlet Poisson, `iters` CG iterations, every launch group bracketed by CUDA
    events on the launching stream (csrc/api.cu profile_cg).  Times are averages per iteration."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import L_lower_equation, Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    shape = [n, n, n] if isinstance(n, int) else list(n)
    mesh = Mesh(Box[0:1, 0:1, 0:1], None, shape, device, dtype)
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(device=device, dtype=var().dtype)
    solver = Solver({"fdm": {"method": "cg", "tol": 1e-30, "max_it": iters, "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == rhs)
    x = var()
    code = N.dtype_code(x.dtype)
    grid = L.lower_grid(mesh.nx, var.bcs)
    faces, nfaces, keep = L.lower_faces(var.bcs, mesh.grid, x, 0, 3)
    eq, keep_e = L_lower_equation(solver.eqs, var)
    lib = N.lib()
    wsb = lib.pa_solver_workspace_bytes(grid, code, N.METHOD["cg"])
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    x_alt = torch.empty_like(x)
    out = (C.c_double * 6)()
    # warm-up pass, then the measured pass
    for _ in range(2):
        x.zero_()
        N.check(lib.pa_cg_profile(grid, eq, nfaces, faces, code, x.data_ptr(), x_alt.data_ptr(), rhs.data_ptr(),
                                  iters, variant, ws.data_ptr(), wsb, out, N.current_stream(x.device)))
    a, b, c, tot, launches, tiled = list(out)
    names = {2.0: ("k_cg_phaseA_tma<double,2>", "k_cg_phaseB_tma<double,2>"),
             1.0: ("k_cg_phaseA<double,2>", "k_cg_phaseB<double,2>"),
             0.0: ("k_cg_dupdate+k_cg_dAd", "k_cg_update")}[tiled]
    return {"phaseA_ms": a, "phaseB_ms": b, "small_ms": c, "iter_ms": tot, "launches_per_iter": launches,
            "tiled": bool(tiled), "path": {2.0: "tma", 1.0: "register-tiled", 0.0: "generic"}[tiled],
            "kernels": names, "share": {"phaseA": a / tot, "phaseB": b / tot, "bc+shell": c / tot}}


def operator_apply_times(shape, op: str = "laplacian", dtype: str = "double", reps: int = 20, variant: str = "auto",
                         kinds=None, vals=None, device: str = "cuda") -> dict:
    """Device time of ONE explicit operator application (`pa_stencil_apply` / `pa_grad_apply` through the C
    ABI, exactly what `Solver.Aop` / `FDC().laplacian/.grad/.div` call), CUDA events on the launching stream
    around `reps` back-to-back calls after 3 warm-up calls.  Calls rotate over enough distinct input/output
    buffer pairs to exceed 2x the 126 MB L2, so every call reads its input from HBM.
    op: "laplacian" | "grad" | "div_upwind" | "div_central" | "advdiff" (div + laplacian, 2 operators).
    Algorithmic words per cell: 2 (R phi, W out), Grad 1 + ndim."""
    import math
    import os

    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdc import FDC
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import L_lower_equation
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    shape = list(shape)
    nd = len(shape)
    kinds = kinds or ["dirichlet"] * (2 * nd)
    vals = vals or [0.0] * (2 * nd)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, device, dtype)
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    x = var()
    esz = x.element_size()
    cells = x.numel()
    ncomp = nd if op == "grad" else 1
    pair_bytes = cells * esz * (1 + ncomp)
    nbuf = max(1, min(64, math.ceil(2.2 * 126e6 / pair_bytes)))
    g = torch.Generator(device=device).manual_seed(7)
    ins = [torch.rand(x.shape, generator=g, dtype=x.dtype, device=device) - 0.5 for _ in range(nbuf)]
    outs = [torch.empty((1, ncomp, *shape) if op == "grad" else x.shape, dtype=x.dtype, device=device)
            for _ in range(nbuf)]
    code, lib = N.dtype_code(x.dtype), N.lib()
    grid = L.lower_grid(mesh.nx, var.bcs)
    stream = N.current_stream(x.device)
    if op == "grad":
        coeffs = FDC({"grad": {"edge": False}}).grad.build_A_coeffs(var)
        popr, keep = L.lower_op(coeffs, nd, x.dtype, dx=mesh._dx, field_device=x.device)

        def call(i):
            N.check(lib.pa_grad_apply(grid, popr, code, ins[i].data_ptr(), outs[i].data_ptr(), stream))
    else:
        fdm = FDM({"div": {"limiter": "none" if op == "div_central" else "upwind", "edge": False}})
        term = {"laplacian": lambda: fdm.laplacian(1.0, var), "div_upwind": lambda: fdm.div(1.0, var),
                "div_central": lambda: fdm.div(1.0, var),
                "advdiff": lambda: fdm.div(1.0, var) - fdm.laplacian(0.1, var)}[op]()
        eq, keep = L_lower_equation(term.ops, var)
        term.ops = {}

        def call(i):
            N.check(lib.pa_stencil_apply(grid, eq, code, ins[i].data_ptr(), outs[i].data_ptr(), stream))

    # "auto": the library's choice (direct 2-D kernel up to 8 M cells, TMA star engine elsewhere); "tma" / "generic":
    # forced paths (PA_APPLY_VARIANT, read per call)
    prev = os.environ.get("PA_APPLY_VARIANT")
    if variant in ("generic", "tma"):
        os.environ["PA_APPLY_VARIANT"] = variant
    else:
        os.environ.pop("PA_APPLY_VARIANT", None)
    try:
        for i in range(3):
            call(i % nbuf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            call(i % nbuf)
        e1.record()
        torch.cuda.synchronize()
    finally:
        if prev is None:
            os.environ.pop("PA_APPLY_VARIANT", None)
        else:
            os.environ["PA_APPLY_VARIANT"] = prev
    ms = e0.elapsed_time(e1) / reps
    words = 1 + ncomp
    return {"op": op, "shape": shape, "dtype": dtype, "variant": variant, "ms": ms, "reps": reps, "buffers": nbuf,
            "GLUP/s": cells / (ms * 1e-3) / 1e9, "words_per_cell": words,
            "GB/s": cells * words * esz / (ms * 1e-3) / 1e9,
            "l2_policy": f"{nbuf} rotating buffer pairs, {nbuf * pair_bytes / 2**20:.0f} MiB > 2 x L2"}


# ---------------------------------------------------------------------------------------------
# The other BASELINE.json configs as timed runs (bench.py `secondary`, tools/bench_configs.py)
# ---------------------------------------------------------------------------------------------
MIXED_BCS = (["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"], [None, None, 0.5, None, 0.0, 0.0])


def _event_time(fn, reps: int = 1, warm: int = 1) -> float:
    """Device ms of `reps` calls of fn after `warm` warm-up calls (CUDA events, current stream).  Solves that last only
    a few ms take warm > 1: a launch that starts on an idle-clocked GPU is over before the clocks are up."""
    for _ in range(max(1, warm)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def solver_throughput(shape, method: str, iters: int, kinds=None, vals=None, dtype: str = "double", variant: int = 0,
                      device: str = "cuda", reps: int = 1, contract: bool = False, warm: int = 1) -> dict:
    """Fixed-count solve (tol 1e-300) of the Poisson problem through the public API; GLUP/s = cells x
    iterations / device time of solver.solve().  words per LUP: SURVEY.md §8d (CG 8, BiCGSTAB 17 canonical,
    Jacobi 3)."""
    import warnings

    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    shape = list(shape)
    nd = len(shape)
    kinds = kinds or ["dirichlet"] * (2 * nd)
    vals = vals or [0.0] * (2 * nd)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, device, dtype)
    tdt = torch.float64 if dtype == "double" else torch.float32
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(tdt).to(device)
    work = rhs.clone()
    max_it = iters - 1 if method in ("cg", "jacobi") else iters
    launches = [0]

    def run():
        work.copy_(rhs)  # set_eq adds the Neumann adjustment in place
        var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
        s = Solver({"fdm": {"method": method, "tol": 1e-300, "max_it": max_it, "report": False, "variant": variant,
                            "check_every": iters + (iters & 1), "contract": contract}})
        s.set_eq(FDM().laplacian(1.0, var) == work)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rep = s.solve()
        assert rep["itr"] == iters, rep
        launches[0] = getattr(var, "_last_launches", 0)

    ms = _event_time(run, reps, warm)
    torch.set_default_dtype(torch.float64)
    cells = 1
    for v in shape:
        cells *= v
    words = {"cg": 8, "bicgstab": 17, "jacobi": 3}[method]
    esz = 8 if dtype == "double" else 4
    glups = cells * iters / (ms * 1e-3) / 1e9
    return {"shape": shape, "method": method, "dtype": dtype, "iters": iters, "ms": ms, "GLUP/s": glups,
            "words_per_lup": words, "GB/s": glups * words * esz, "launches": launches[0]}


def euler_throughput(shape, limiter: str, steps: int, dtype: str = "double", device: str = "cuda", warm: int = 1) -> dict:
    """Config 3: explicit Euler steps of the advection-diffusion equation ddt + div(u phi) - nu lap(phi) = 0,
    u = 1, nu = 0.1, Dirichlet 0, phi0 = rand(seed 1234); 2 words per LUP (R phi, W phi_new)."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    shape = list(shape)
    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, device, dtype)
    var = Field("c", 1, mesh, {"domain": homogeneous_bcs(nd, 0.0, "dirichlet"), "obstacle": None})
    g = torch.Generator().manual_seed(1234)
    var.set_var_tensor(torch.rand(1, *shape, generator=g, dtype=torch.float64).to(var().dtype).to(device))
    nu, u = 0.1, 1.0
    # SURVEY §8d's dt = 0.2 dx^2/nu is beyond the explicit diffusion limit dx^2/(2 nd nu) in 3-D; half
    # the limit keeps a long run finite
    var.set_time(0.5 * min(mesh._dx) ** 2 / (2 * nd * nu), 0.0)
    fdm = FDM({"div": {"limiter": limiter, "edge": False}})
    s = Solver({"fdm": {"method": "euler", "tol": 0.0, "max_it": 0, "report": False, "n_steps": steps}})
    s.set_eq(fdm.ddt(var) + fdm.div(u, var) - fdm.laplacian(nu, var) == 0.0)
    ms = _event_time(lambda: s.solve(), 1, warm)
    torch.set_default_dtype(torch.float64)
    cells = 1
    for v in shape:
        cells *= v
    esz = 8 if dtype == "double" else 4
    glups = cells * steps / (ms * 1e-3) / 1e9
    assert bool(torch.isfinite(var()).all())
    return {"shape": shape, "method": f"euler/{limiter}", "dtype": dtype, "iters": steps, "ms": ms, "GLUP/s": glups,
            "words_per_lup": 2, "GB/s": glups * 2 * esz}
