"""Analytic fixtures of the Poisson test problems (reference: pyapes/testing/poisson.py).
Host-side helpers used by tests and demos."""
from math import pi

import torch
from torch import Tensor

from pyapes_b200.geometry.basis import FDIR
from pyapes_b200.mesh import Mesh
from pyapes_b200.variables import Field


def poisson_rhs_nd(mesh: Mesh, var: Field) -> Tensor:
    rhs = torch.zeros_like(var())
    if mesh.dim == 1:
        rhs[0] = 1.0 - 2.0 * mesh.X**2
    elif mesh.dim == 2:
        rhs[0] = 6.0 * mesh.X * mesh.Y * (1.0 - mesh.Y) - 2.0 * (mesh.X**3)
    else:
        rhs[0] = torch.sin(pi * mesh.X) * torch.sin(pi * mesh.Y) * torch.sin(pi * mesh.Z)
    return rhs


def poisson_exact_nd(mesh: Mesh) -> Tensor:
    if mesh.dim == 1:
        return 7.0 / 9.0 - 2.0 / 9.0 * mesh.X + mesh.X**2 / 2.0 - mesh.X**4 / 6.0
    if mesh.dim == 2:
        return mesh.Y * (1.0 - mesh.Y) * (mesh.X**3)
    return -1.0 / (3 * pi**2) * torch.sin(pi * mesh.X) * torch.sin(pi * mesh.Y) * torch.sin(pi * mesh.Z)


def poisson_1d_bc(grid, mask: Tensor, *_) -> Tensor:
    x = grid[0][mask]
    return 7.0 / 9.0 - 2.0 / 9.0 * x + x**2 / 2.0 - x**4 / 6.0


def poisson_2d_bc(grid, mask: Tensor, *_) -> Tensor:
    return grid[1][mask] * (1.0 - grid[1][mask]) * (grid[0][mask] ** 3)


def poisson_bcs(dim: int = 3, debug: bool = False) -> list:
    val = {1: poisson_1d_bc, 2: poisson_2d_bc}.get(dim, 0.0)
    return [{"bc_face": FDIR[i], "bc_type": "dirichlet", "bc_val": 4.44 if debug else val, "bc_val_opt": None}
            for i in range(dim * 2)]
