"""2-D (single-plane) problems: TMA pipeline vs generic kernels."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from tools.bench_configs import solver_case
D4 = (["dirichlet"] * 4, [0.0] * 4)
for n in (64, 256, 1024, 4096):
    for variant in (0, 1, 2):
        for method, iters, words in (("cg", 200, 8), ("jacobi", 200, 3)):
            if method == "jacobi" and variant == 2:
                continue
            print(json.dumps(solver_case(f"{method} {n}^2 v{variant}", [n, n], method, D4, iters, words, variant=variant)))
