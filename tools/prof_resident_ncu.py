"""One resident Euler launch (1024^2, 200 steps) and one resident CG solve for an ncu capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

P.euler_throughput([1024, 1024], "upwind", 200)
P.solver_throughput([1024, 1024], "cg", 100, variant=6)
