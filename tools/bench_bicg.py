"""BiCGSTAB / Jacobi / Euler throughput at 512^3 (subset of tools/bench_configs.py)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_configs as B
import torch
D6 = (["dirichlet"] * 6, [0.0] * 6)
out = [B.solver_case("config4 BiCGSTAB 512^3 mixed", [512] * 3, "bicgstab", B.MIXED, 20, 17),
       B.solver_case("BiCGSTAB 512^3 Dirichlet", [512] * 3, "bicgstab", D6, 20, 17),
       B.solver_case("config4 Jacobi 512^3 mixed", [512] * 3, "jacobi", B.MIXED, 50, 3)]
torch.set_default_dtype(torch.float64)
out.append(B.euler_case("Euler 512^3 upwind_fd", [512] * 3, "upwind_fd", 20))
for o in out:
    print(json.dumps(o))
