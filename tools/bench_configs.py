"""Throughput of the other BASELINE.json configs (2-4) on one GPU, as GLUP/s and fraction of the
HBM roofline for their algorithmic traffic.  Not the headline bench.
usage: python tools/bench_configs.py [--quick]"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.filterwarnings("ignore")
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs, mixed_bcs

HBM = 6541.8e9
quick = "--quick" in sys.argv
MIXED = (["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"], [None, None, 0.5, None, 0.0, 0.0])


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def solver_case(name, shape, method, bcs, iters, words, variant=0, dtype="double"):
    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda", dtype)
    esz = 8 if dtype == "double" else 4
    kinds, vals = bcs
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(mesh.dtype.float).cuda()
    max_it = iters - 1 if method in ("cg", "jacobi") else iters

    def run():
        var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
        s = Solver({"fdm": {"method": method, "tol": 1e-300, "max_it": max_it, "report": False, "variant": variant,
                            "check_every": iters + (iters & 1)}})
        s.set_eq(FDM().laplacian(1.0, var) == rhs.clone())
        rep = s.solve()
        assert rep["itr"] == iters, rep
    ms = timed(run)
    n = 1
    for v in shape: n *= v
    glups = n * iters / (ms * 1e-3) / 1e9
    return {"case": name, "shape": shape, "method": method, "iters": iters, "ms": ms, "GLUP/s": round(glups, 2),
            "dtype": dtype, "words_per_LUP": words, "hbm_frac": round(glups * 1e9 * words * esz / HBM, 3), "variant": variant}


def euler_case(name, shape, limiter, steps):
    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda")
    var = Field("c", 1, mesh, {"domain": homogeneous_bcs(nd, 0.0, "dirichlet"), "obstacle": None})
    g = torch.Generator().manual_seed(1234)
    var.set_var_tensor(torch.rand(1, *shape, generator=g, dtype=torch.float64).cuda())
    nu, u = 0.1, 1.0
    # SURVEY §8d asks for dt = 0.2 dx^2/nu, which is beyond the explicit diffusion limit dx^2/(2 nd nu)
    # in 3-D (the field overflows after ~10^3 steps); half the limit keeps the run finite
    var.set_time(0.5 * min(mesh._dx) ** 2 / (2 * nd * nu), 0.0)
    fdm = FDM({"div": {"limiter": limiter, "edge": False}})
    s = Solver({"fdm": {"method": "euler", "tol": 0.0, "max_it": 0, "report": False, "n_steps": steps}})
    s.set_eq(fdm.ddt(var) + fdm.div(u, var) - fdm.laplacian(nu, var) == 0.0)
    ms = timed(lambda: s.solve())
    n = 1
    for v in shape: n *= v
    glups = n * steps / (ms * 1e-3) / 1e9
    return {"case": name, "shape": shape, "method": f"euler/{limiter}", "iters": steps, "ms": ms, "GLUP/s": round(glups, 2),
            "words_per_LUP": 2, "hbm_frac": round(glups * 1e9 * 16 / HBM, 3)}


if __name__ == "__main__":
    out = []
    N3 = 256 if quick else 512
    D6 = (["dirichlet"] * 6, [0.0] * 6)
    out.append(solver_case("config2 CG 256^3 Dirichlet", [256] * 3, "cg", D6, 50, 8))
    out.append(solver_case(f"config4 BiCGSTAB {N3}^3 mixed", [N3] * 3, "bicgstab", MIXED, 20, 17))
    out.append(solver_case(f"config4 Jacobi {N3}^3 mixed", [N3] * 3, "jacobi", MIXED, 50, 3))
    out.append(solver_case(f"Jacobi {N3}^3 Dirichlet", [N3] * 3, "jacobi", D6, 50, 3))
    out.append(solver_case(f"BiCGSTAB {N3}^3 Dirichlet", [N3] * 3, "bicgstab", D6, 20, 17))
    out.append(solver_case(f"BiCGSTAB {N3}^3 Dirichlet generic", [N3] * 3, "bicgstab", D6, 20, 17, variant=1))
    out.append(solver_case("CG 512^3 Dirichlet fp32", [512] * 3, "cg", D6, 50, 8, dtype="single"))
    out.append(solver_case("CG 1024^2 Dirichlet (2-D)", [1024, 1024], "cg", (["dirichlet"] * 4, [0.0] * 4), 200, 8))
    torch.set_default_dtype(torch.float64)
    out.append(euler_case("config3 Euler 256^3 upwind", [256] * 3, "upwind", 100))
    out.append(euler_case("config3 Euler 256^3 upwind_fd", [256] * 3, "upwind_fd", 100))
    out.append(euler_case("config3 Euler 1024^2 upwind", [1024, 1024], "upwind", 100))
    out.append(euler_case(f"Euler {N3}^3 upwind", [N3] * 3, "upwind", 20))
    for o in out:
        print(json.dumps(o))
