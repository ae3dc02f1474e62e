"""One short run of every kernel family for a multi-kernel ncu capture (profiles/r02_ncu_full_kernels.txt):
BiCGSTAB / Jacobi 512^3 with config-4 BCs (k_star_tma APPLY_V / JACOBI, k_bi_st_tma, k_bi_x_stream, k_bc_face_pair,
k_shell_norm), the small-grid whole-solve kernels (k_cg_persistent 32^3, k_cg_coop_tma 96^3), the resident CG
(1024^2) and the explicit operators (Laplacian / Grad 512^3).  usage: python tools/prof_kernels.py"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import pyapes_b200.profile as P  # noqa: E402

P.solver_throughput([512] * 3, "bicgstab", 2, *P.MIXED_BCS)
P.solver_throughput([512] * 3, "jacobi", 2, *P.MIXED_BCS)
P.solver_throughput([32] * 3, "cg", 6, variant=3)
P.solver_throughput([96] * 3, "cg", 6, variant=5)
P.solver_throughput([1024, 1024], "cg", 6, variant=6)
P.operator_apply_times([512] * 3, "laplacian", reps=1)
P.operator_apply_times([512] * 3, "grad", reps=1)
