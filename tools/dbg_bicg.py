import sys, warnings; sys.path.insert(0, '/root/repo')
import torch
from tests import _util as U
from tests.test_gpu_parity import _run_solver_case
SOL = U.load("solvers.pt")
for case in SOL:
    if case["method"] != "bicgstab": continue
    for variant in (1,):
        var, rep, w = _run_solver_case(case, variant=variant)
        ref = case["report"]
        sol = var().cpu()
        err = (sol - case["solution"]).abs().max().item() if "solution" in case else float('nan')
        print(f"{case['name']:32s} gpu itr={rep['itr']:5d} tol={rep['tol']:.6e} | ref itr={ref['itr']:5d} tol={ref['tol']:.6e} | max|dsol|={err:.3e}")
