"""Callable boundary values shared by the fixture generator (run against the real reference) and
the GPU tests (run against pyapes_b200): signature `(grid, mask, var, opt)` (bcs.py:32,203-205)."""
import torch


# Polynomials only: +, -, * are correctly rounded on the CPU (where the fixtures are made) and on the GPU
# (where the tests evaluate the callables), transcendental functions are not bit-identical between the two.
def neumann_cos(grid, mask, *_):
    y = grid[1][mask]
    return 0.3 * (1.0 - 2.0 * y * y)


def dirichlet_sin(grid, mask, *_):
    y = grid[1][mask]
    return 3.0 * y * (1.5 - y) + grid[0][mask]


def dirichlet_of_var(grid, mask, var, *_):
    """Depends on the field itself: the reference re-evaluates it at every BC application."""
    return 0.5 * var[0][mask] + 1.0


CALLABLES = {"neumann_cos": neumann_cos, "dirichlet_sin": dirichlet_sin, "dirichlet_of_var": dirichlet_of_var}
