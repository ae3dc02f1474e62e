"""How far does the reference's own BiCGSTAB iteration count move under a 1-ulp-level
perturbation of the RHS?  (CPU, oracle == reference bit for bit.)  Used to set the parity band of
the GPU BiCGSTAB tests; results are quoted in DESIGN.md."""
import sys, warnings
sys.path.insert(0, '/root/repo')
import torch
from oracle import fd_oracle as O
from tests import _util as U
warnings.filterwarnings("ignore")
SOL = U.load("solvers.pt")
for case in SOL:
    if case["method"] != "bicgstab" or case["spec"]["dtype"] != "double": continue
    xs, dx = U.oracle_axes(case); bcs = U.oracle_bcs(case)
    shape = (1, *case["spec"]["nx"])
    its, sols = [], []
    for trial in range(6):
        x = torch.zeros(shape, dtype=torch.float64) + case["init"]
        rhs = U.case_rhs(case, shape, torch.float64)
        if trial:
            g = torch.Generator().manual_seed(trial)
            rhs = rhs * (1 + 2.2e-16 * torch.randn(shape, generator=g, dtype=torch.float64))
        eq = O.Equation(U.oracle_terms(case), dx, xs, bcs).build(x)
        eq.adjust_rhs(x, rhs)
        sol, rep, _ = O.bicgstab(eq, x, rhs, case["tol"], case["max_it"])
        its.append(rep["itr"]); sols.append(sol)
    dmax = max((s - sols[0]).abs().max().item() for s in sols[1:])
    print(f"{case['name']:30s} itr unperturbed={its[0]:4d} perturbed={its[1:]}  max|dsol|={dmax:.2e}")
