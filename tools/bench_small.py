"""CG on small grids: per-iteration latency of the kernel variants (0 auto, 3 persistent generic kernel,
4 fused TMA kernels as separate launches, 5 cooperative whole-solve TMA kernel)."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.filterwarnings("ignore")
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs

def case(shape, variant, iters=100, dtype="double"):
    nd = len(shape)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda", dtype)
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).to(mesh.dtype.float).cuda()
    res = {}
    def run():
        var = Field("p", 1, mesh, {"domain": homogeneous_bcs(nd, 0.0, "dirichlet"), "obstacle": None})
        s = Solver({"fdm": {"method": "cg", "tol": 1e-300, "max_it": iters - 1, "report": False, "variant": variant,
                            "check_every": iters}})
        s.set_eq(FDM().laplacian(1.0, var) == rhs.clone())
        rep = s.solve()
        assert rep["itr"] == iters, rep
        res["tol"] = rep["tol"]; res["sum"] = var().sum().item()
    run(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = 1
    for v in shape: n *= v
    return {"shape": shape, "variant": variant, "us_per_iter": round(best / iters * 1e3, 2),
            "GLUP/s": round(n * iters / best / 1e6, 2), "tol": res["tol"], "sum": res["sum"]}

for shape in ([64, 64], [128, 128], [256, 256], [512, 512], [1024, 1024], [1448, 1448], [32, 32, 32], [64, 64, 64], [96, 96, 96], [128, 128, 128]):
    for v in (3, 4, 5):
        if v == 3 and shape[0] ** len(shape) > 300000:
            continue
        a, b = case(shape, v, 60), case(shape, v, 180)
        pure = (b["us_per_iter"] * 180 - a["us_per_iter"] * 60) / 120  # setup cancels
        print(json.dumps({"shape": shape, "variant": v, "us_per_iter_180": b["us_per_iter"], "us_per_iter_marginal": round(pure, 2)}), flush=True)
