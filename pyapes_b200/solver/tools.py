"""Solver configuration types (reference: pyapes/solver/tools.py:13-26).

`default_A_ops` — the reference's full-size coefficient-tensor factory (tools.py:29-112) — has no
counterpart: constant coefficients are three numbers per axis here (pyapes_b200/_lower.py)."""
from typing import TypedDict


class FDMSolverConfig(TypedDict, total=False):
    method: str   # "cg" | "bicgstab" | "jacobi" (jacobi is new)
    tol: float
    max_it: int
    report: bool
    # optional knobs of this implementation (ignored by the reference)
    check_every: int  # iterations between host polls of the device-side convergence latch
    use_graph: bool   # replay the iteration as a CUDA graph
    variant: int      # 0 auto, 1 generic kernels, 2 register-tiled, 3 persistent small-grid CG, 4 fused (TMA) only
    n_steps: int      # explicit Euler: time steps per solve()


class SolverConfig(TypedDict):
    fdm: FDMSolverConfig
