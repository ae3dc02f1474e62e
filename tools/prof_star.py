"""Short runs of the non-CG paths for ncu: BiCGSTAB / Jacobi (mixed BCs of config 4) and the
explicit Euler adv-diff step (config 3) at n^3.  usage: python tools/prof_star.py [n] [what...]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs, mixed_bcs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
what = sys.argv[2:] or ["bicgstab", "jacobi", "euler"]
mesh = Mesh(Box[0:1, 0:1, 0:1], None, [n] * 3, "cuda")
g = torch.Generator().manual_seed(1234)
rhs = torch.rand(1, n, n, n, generator=g, dtype=torch.float64).cuda()
MIXED = mixed_bcs([None, None, 0.5, None, 0.0, 0.0], ["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"])
for method in ("bicgstab", "jacobi"):
    if method not in what:
        continue
    var = Field("p", 1, mesh, {"domain": MIXED, "obstacle": None})
    s = Solver({"fdm": {"method": method, "tol": 1e-300, "max_it": 4, "report": False, "use_graph": False}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs.clone())
    print(method, s.solve())
if "euler" in what:
    var = Field("c", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    var.set_var_tensor(rhs.clone())
    var.set_time(0.05 * min(mesh._dx) ** 2 / 0.1, 0.0)
    fdm = FDM({"div": {"limiter": "upwind_fd", "edge": False}})
    s = Solver({"fdm": {"method": "euler", "report": False, "n_steps": 3}})
    s.set_eq(fdm.ddt(var) + fdm.div(1.0, var) - fdm.laplacian(0.1, var) == 0.0)
    s.solve()
    torch.cuda.synchronize()
    print("euler ok")
