"""Solver configuration types (reference: pyapes/solver/tools.py:13-26) and `default_A_ops`.

The kernels never read coefficient tensors: constant coefficients are three numbers per axis
(pyapes_b200/_lower.py).  `default_A_ops` is kept for user code that inspects the reference's
full-size `[App, Ap, Ac, Am, Amm]` lists (tools.py:29-112); nothing in this package calls it."""
from typing import TypedDict

import torch
from torch import Tensor

# second-order central stars: (phi[+1], phi[0], phi[-1]) weights before the 1/dx scaling
_STAR = {"grad": (1.0, 0.0, -1.0), "div": (1.0, 0.0, -1.0), "laplacian": (1.0, -2.0, 1.0)}


def default_A_ops(var, ops: str) -> list[list[Tensor]]:
    """`[App, Ap, Ac, Am, Amm]`, each a list over mesh axes of tensors shaped like `var()`: the
    unedited coefficients of `ops` in {"grad", "div", "laplacian"} for neighbours i+2 .. i-2
    (tools.py:29-112).  Axisymmetric meshes scale the r axis: Laplacian `Ap, Am = 1 +- dr/(2r)`,
    Div `Ac = 2 dr / r`, with 0 where r = 0."""
    key = ops.lower()
    if key not in _STAR:
        raise RuntimeError(f"Given {ops=} should be either grad, div, or laplacian.")
    up, mid, dn = _STAR[key]
    one, nd = torch.ones_like(var()), var.mesh.dim
    rz = var.mesh.coord_sys == "rz" and key != "grad"
    if rz:
        ratio = var.mesh.dx[0] / var.mesh.R
        bend = torch.nan_to_num(ratio / 2 if key == "laplacian" else 2 * ratio, nan=0.0, posinf=0.0, neginf=0.0)
    out: list[list[Tensor]] = [[], [], [], [], []]
    for axis in range(nd):
        w = [0.0 * one, up * one, mid * one, dn * one, 0.0 * one]
        if rz and axis == 0:
            if key == "laplacian":
                w[1], w[3] = (1 + bend) * one, (1 - bend) * one
            else:
                w[2] = bend * one
        for slot, t in zip(out, w):
            slot.append(t)
    return out


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
    contract: bool    # opt-in FMA contraction in the fused TMA CG kernels (PA_FLAG_CONTRACT): ~1e-16 relative per
    #                   operation instead of bit-exact, fewer fp64 instructions; default False


class SolverConfig(TypedDict):
    fdm: FDMSolverConfig
