"""Exact solution of the 1-D viscous Burgers problem used by the reference's work-in-progress
nonlinear-advection test (reference: pyapes/testing/burgers.py; tests/test_solver.py:393-436).
Host-side helper."""
from math import pi

import torch
from torch import Tensor

from pyapes_b200.mesh import Mesh


def burger_exact_nd(mesh: Mesh, nu: float, t: float) -> Tensor:
    """Cole-Hopf solution `u = -2 nu phi_x / phi + 4` with `phi` the sum of two Gaussians that
    travel at speed 4 and spread like `4 nu (t + 1)`; 1-D only (2-D/3-D raise like the reference)."""
    if mesh.dim != 1:
        raise NotImplementedError
    spread = 4 * nu * (t + 1)
    phi = torch.zeros_like(mesh.X)
    dphi_dx = torch.zeros_like(mesh.X)
    for shift in (0.0, 2 * pi):
        xi = mesh.X - 4 * t - shift
        bump = torch.exp(-(xi**2) / spread)
        phi = phi + bump
        dphi_dx = dphi_dx - 0.5 * xi / (nu * (t + 1)) * bump
    return -2 * nu * dphi_dx / phi + 4
