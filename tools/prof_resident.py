"""Phase time stamps of the resident kernels (PA_RES_DEBUG): python tools/prof_resident.py"""
import os
import sys

os.environ["PA_RES_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

for n in (256, 1024):
    print(f"--- {n}^2", file=sys.stderr, flush=True)
    P.euler_throughput([n, n], "upwind", 40)
    P.solver_throughput([n, n], "cg", 40, variant=6)
