"""One short run of a kernel family for an ncu capture (profiles/r02_ncu_kernels_*.txt).
usage: python tools/prof_kernels.py bicgstab|jacobi|small|coop|resident_cg|apply"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch  # noqa: E402

from pyapes_b200.geometry import Box  # noqa: E402
from pyapes_b200.mesh import Mesh  # noqa: E402
from pyapes_b200.solver.fdm import FDM  # noqa: E402
from pyapes_b200.solver.ops import Solver  # noqa: E402
from pyapes_b200.variables import Field  # noqa: E402
from pyapes_b200.variables.bcs import mixed_bcs  # noqa: E402
import pyapes_b200.profile as P  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "jacobi"


def solve(shape, method, iters, kinds=None, vals=None, variant=0):
    nd = len(shape)
    kinds = kinds or ["dirichlet"] * (2 * nd)
    vals = vals or [0.0] * (2 * nd)
    mesh = Mesh(Box([0.0] * nd, [1.0] * nd), None, shape, "cuda", "double")
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, *shape, generator=g, dtype=torch.float64).cuda()
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    s = Solver({"fdm": {"method": method, "tol": 1e-300, "max_it": iters, "report": False, "variant": variant,
                        "use_graph": False}})
    s.set_eq(FDM().laplacian(1.0, var) == rhs)
    print(what, s.solve())


if what == "bicgstab":
    solve([512] * 3, "bicgstab", 2, *P.MIXED_BCS)
elif what == "jacobi":
    solve([512] * 3, "jacobi", 1, *P.MIXED_BCS)
elif what == "small":
    solve([32] * 3, "cg", 20, variant=3)
elif what == "coop":
    solve([96] * 3, "cg", 20, variant=5)
elif what == "resident_cg":
    solve([1024, 1024], "cg", 100, variant=6)
elif what == "apply":
    print(P.operator_apply_times([512] * 3, "laplacian", reps=1)["ms"], P.operator_apply_times([512] * 3, "grad", reps=1)["ms"])
