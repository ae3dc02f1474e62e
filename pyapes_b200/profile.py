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
