"""Fixed cost of one `solver.solve()` call: wall-clock (host) and device time of CG solves with 2, 50 and
200 iterations at several sizes; the intercept of time vs iterations is what a short solve (implicit
Euler steps, the reference's own test problems) pays per call.
usage: python tools/solve_overhead.py"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.filterwarnings("ignore")
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs


def case(shape, iters, reps=7):
    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda", "double")
    rhs = torch.rand(1, *shape, generator=torch.Generator().manual_seed(1234), dtype=torch.float64).cuda()
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(nd, 0.0, "dirichlet"), "obstacle": None})
    s = Solver({"fdm": {"method": "cg", "tol": 1e-300, "max_it": iters - 1, "report": False}})
    wall, dev = [], []
    for _ in range(reps + 2):
        var.set_var_tensor(torch.zeros_like(var()))
        s.set_eq(FDM().laplacian(1.0, var) == rhs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        rep = s.solve()
        e1.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(e0.elapsed_time(e1))
        assert rep["itr"] == iters, rep
    wall, dev = sorted(wall[2:]), sorted(dev[2:])
    return {"shape": shape, "iters": iters, "wall_ms": round(wall[len(wall) // 2], 4), "device_ms": round(dev[len(dev) // 2], 4)}


if __name__ == "__main__":
    for shape in ([64, 64], [256, 256], [64, 64, 64], [256, 256, 256]):
        rows = [case(shape, it) for it in (2, 50, 200)]
        per_it = (rows[2]["wall_ms"] - rows[1]["wall_ms"]) / 150.0
        fixed = rows[1]["wall_ms"] - 50 * per_it
        for r in rows:
            print(json.dumps(r))
        print(json.dumps({"shape": shape, "per_iteration_ms": round(per_it, 5), "fixed_ms_per_solve": round(fixed, 4)}))
