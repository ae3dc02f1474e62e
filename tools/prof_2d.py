"""Short 2-D runs for ncu: CG / Jacobi / Euler on n^2.  usage: python tools/prof_2d.py [n]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mesh = Mesh(Box[0:1, 0:1], None, [n, n], "cuda")
g = torch.Generator().manual_seed(1234)
rhs = torch.rand(1, n, n, generator=g, dtype=torch.float64).cuda()
for method in ("cg", "jacobi"):
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(2, 0.0, "dirichlet"), "obstacle": None})
    s = Solver({"fdm": {"method": method, "tol": 1e-300, "max_it": 6, "report": False, "use_graph": False}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs.clone())
    print(method, s.solve())
var = Field("c", 1, mesh, {"domain": homogeneous_bcs(2, 0.0, "dirichlet"), "obstacle": None})
var.set_var_tensor(rhs.clone())
var.set_time(0.1 * min(mesh._dx) ** 2 / 0.1, 0.0)
fdm = FDM({"div": {"limiter": "upwind_fd", "edge": False}})
s = Solver({"fdm": {"method": "euler", "report": False, "n_steps": 3}})
s.set_eq(fdm.ddt(var) + fdm.div(1.0, var) - fdm.laplacian(0.1, var) == 0.0)
s.solve(); torch.cuda.synchronize(); print("euler ok")
