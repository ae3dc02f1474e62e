"""CPU oracle for the pyapes finite-difference hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product (`pyapes_b200/`) never does: it has no CPU path.

What it is: a torch-CPU restatement of the reference's algorithm (roll-based 5-coefficient
stencils against full-size coefficient tensors, sequential per-face boundary writes,
matrix-free CG / BiCGSTAB), operating on plain tensors instead of the reference's
Field/Mesh objects.  It keeps the reference's *operation order* so that, on the same inputs,
it is bit-identical to the reference on CPU.  That is pinned by `tests/test_oracle_golden.py`
against fixtures produced by the real reference (`tests/golden/make_golden.py`).
Parity status: PINNED for Laplacian/Grad/Div apply, rhs adjustment, BC application, CG and
BiCGSTAB, edge=True one-sided stencils, jacobian/hessian, the axisymmetric (rz) operators and
nonlinear advection `div(var, var)` (tests/golden/{ops,solvers,edges,rz_ops,nonlinear}.pt, all
produced by the real reference).  `jacobi`, `euler_step` and `implicit_euler_step` have NO
counterpart in the reference (fdm.Ddt is a stub): they are defined here from
reference primitives and are "parity unpinned" (the implicit step is checked against a dense solve).

Citations are `file:line` in /root/reference (pyapes v0.2.13).

Conventions
-----------
phi  : tensor (1, *nx) (scalar field; the reference's Krylov solvers only work for dim 1)
dx   : list[float], one per mesh axis           (mesh/_mesh.py:67-77)
xs   : list of 1-D coordinate tensors           (mesh/_mesh.py:82-92)
bcs  : list of FaceBC in application order      (variables/bcs.py:363-440)
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Any

import torch
from torch import Tensor

AXIS_OF = {"x": 0, "y": 1, "z": 2}
AXIS_OF_RZ = {"r": 0, "z": 1}  # geometry/basis.py:10
FACES = ["xl", "xu", "yl", "yu", "zl", "zu"]  # geometry/basis.py:16


@dataclass
class FaceBC:
    """One boundary face.  `value` is a float, None, or a tensor with one entry per face
    cell in row-major order of the remaining axes (what `var[mask]` enumerates)."""

    face: str
    kind: str  # dirichlet | neumann | symmetry | periodic
    value: Any = None
    rz: bool = False

    @property
    def axis(self) -> int:  # bcs.py:72-75 (an "r" face only exists on rz meshes; their "z" is axis 1)
        return AXIS_OF[self.face[0]] if not self.rz else AXIS_OF_RZ[self.face[0]]

    @property
    def side(self) -> int:  # bcs.py:77-80  (-1 lower, +1 upper)
        return -1 if self.face[-1] == "l" else 1


def make_axes(lower, upper, nx, dtype=torch.float64):
    """Node coordinates and spacing for an int-spacing mesh (mesh/_mesh.py:53-92)."""
    lo = torch.tensor([float(v) for v in lower], dtype=dtype)
    up = torch.tensor([float(v) for v in upper], dtype=dtype)
    lx = up - lo
    dx = [float(l / (n - 1.0)) for l, n in zip(lx, nx)]
    xs = [
        torch.linspace(lo[i].item(), up[i].item(), int(nx[i]), dtype=dtype)
        for i in range(len(nx))
    ]
    return xs, dx


# ---------------------------------------------------------------------------------------
# plane helpers: the reference addresses planes through full-size bool masks
# (bcs.py:84-93); a plane slice selects the same cells in the same order.
# ---------------------------------------------------------------------------------------
def _plane(ndim: int, axis: int, idx: int):
    s: list[Any] = [slice(None)] * ndim
    s[axis] = idx
    return tuple(s)


def _face_idx(bc: FaceBC, n: int, shift: int) -> int:
    """Index along bc.axis of the plane `shift` cells inward (negative shift = wrapped
    forward planes, as torch.roll of the mask does, bcs.py:84-93)."""
    base = 0 if bc.side < 0 else n - 1
    return (base - bc.side * shift) % n


def _bc_scalar(bc: FaceBC):
    """fdc._return_bc_val (fdc.py:803-817) for non-callable values."""
    if bc.value is None:
        return 0.0
    return bc.value


# ---------------------------------------------------------------------------------------
# boundary-condition application  (bcs.py:197-280, linalg.py:282-299)
# ---------------------------------------------------------------------------------------
def apply_bcs(phi: Tensor, xs: list[Tensor], bcs: list[FaceBC]) -> None:
    nd = phi.dim() - 1
    v = phi[0]
    for bc in bcs:
        a, n = bc.axis, v.shape[bc.axis]
        face = _plane(nd, a, _face_idx(bc, n, 0))
        p1 = _plane(nd, a, _face_idx(bc, n, 1))
        p2 = _plane(nd, a, _face_idx(bc, n, 2))
        if bc.kind == "dirichlet":  # bcs.py:200-213
            assert bc.value is not None
            val = bc.value
            if isinstance(val, Tensor):
                v[face] = val.reshape(v[face].shape) if val.numel() > 1 else val
            else:
                v[face] = float(val)
        elif bc.kind == "neumann":  # bcs.py:223-253
            assert bc.value is not None
            d = xs[a][_face_idx(bc, n, 0)] - xs[a][_face_idx(bc, n, 1)]
            dxv = torch.zeros_like(v[face]) + d  # grid[mask] - grid[mask_prev]
            val = bc.value
            if isinstance(val, Tensor):
                c = val.reshape(v[face].shape) if val.numel() > 1 else val
            else:
                c = float(val)
            v[face] = 4 / 3 * v[p1] - 1 / 3 * v[p2] + 2 / 3 * c * dxv * bc.side
        elif bc.kind == "symmetry":  # bcs.py:259-262
            v[face] = v[p1]
        elif bc.kind == "periodic":  # bcs.py:268-280
            if bc.side < 0:
                f1 = _plane(nd, a, _face_idx(bc, n, -1))
                f2 = _plane(nd, a, _face_idx(bc, n, -2))
                v[face] = v[p1] - v[f1] + v[f2]
            else:
                f1 = _plane(nd, a, _face_idx(bc, n, -1))
                v[face] = v[f1]
        else:
            raise ValueError(bc.kind)


def solver_region(nd: int, bcs: list[FaceBC]):
    """mesh/tools.py:7-20 — [1:-1] per axis, open on periodic sides."""
    se: list[list[Any]] = [[1, -1] for _ in range(nd)]
    for bc in bcs:
        if bc.kind == "periodic":
            se[bc.axis][0 if bc.side < 0 else 1] = None
    return tuple(slice(*p) for p in se)


# ---------------------------------------------------------------------------------------
# coefficient tensors  [App, Ap, Ac, Am, Amm], each a list over mesh axes (tools.py:29-112)
# ---------------------------------------------------------------------------------------
def _defaults(phi: Tensor, ap: float, ac: float, am: float):
    nd = phi.dim() - 1
    z = lambda: [torch.zeros_like(phi) for _ in range(nd)]  # noqa: E731
    f = lambda c: [c * torch.ones_like(phi) for _ in range(nd)]  # noqa: E731
    return z(), f(ap), f(ac), f(am), z()


def _rz_grid(phi: Tensor, xs: list[Tensor]):
    """meshgrid of the (r, z) node coordinates, as Mesh.grid (mesh/_mesh.py:95)."""
    return torch.meshgrid(xs, indexing="ij")


def laplacian_coeffs(phi: Tensor, dx: list[float], bcs: list[FaceBC], rz_xs: list[Tensor] | None = None):
    """fdc.py:375-423; `rz_xs` (node coordinates) switches on the axisymmetric branches
    (tools.py:86-108, fdc.py:395-403): the r-axis coefficients carry 1 +- dr/(2r)."""
    App, Ap, Ac, Am, Amm = _defaults(phi, 1.0, -2.0, 1.0)
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    grid = None
    if rz_xs is not None:
        grid = _rz_grid(phi, rz_xs)
        scale = torch.nan_to_num(dxt[0] / (2 * grid[0]), nan=0.0, posinf=0.0, neginf=0.0)
        Ap[0] = (1 + scale) * torch.ones_like(phi)
        Am[0] = (1 - scale) * torch.ones_like(phi)
    for j in range(nd):
        n = phi.shape[1 + j]
        for bc in bcs:
            if bc.axis != j:
                continue
            if bc.kind in ("neumann", "symmetry"):
                pl = _plane(nd, j, _face_idx(bc, n, 1))
                alpha = torch.zeros_like(Ap[j][0][pl])
                if grid is not None:
                    dr = dxt[j] if j == 0 else 0.0
                    alpha = torch.nan_to_num(2 / 3 * dr / grid[j][pl], nan=0.0, posinf=0.0, neginf=0.0)
                if bc.side < 0:
                    Ap[j][0][pl] = 2 / 3 + alpha
                    Ac[j][0][pl] = -(2 / 3 + alpha)
                    Am[j][0][pl] = 0.0
                else:
                    Ap[j][0][pl] = 0.0
                    Ac[j][0][pl] = -(2 / 3 + alpha)
                    Am[j][0][pl] = 2 / 3 + alpha
        Ap[j][0] /= dxt[j] ** 2
        Ac[j][0] /= dxt[j] ** 2
        Am[j][0] /= dxt[j] ** 2
    return [App, Ap, Ac, Am, Amm]


def _central_edit(phi, dx, bcs, Ap, Ac, Am, gamma_min=None, gamma_max=None):
    """fdc.py:543-609.  gamma_* are (1,*nx) tensors or None (=ones)."""
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    if gamma_min is None:
        gamma_min = torch.ones_like(phi)
        gamma_max = torch.ones_like(phi)
    for j in range(nd):
        n = phi.shape[1 + j]
        for bc in bcs:
            if bc.axis != j:
                continue
            pl = _plane(nd, j, _face_idx(bc, n, 1))
            if bc.kind in ("neumann", "symmetry"):
                gmx = gamma_max[0][pl]
                gmn = gamma_min[0][pl]
                if bc.side < 0:
                    Ap[j][0][pl] += 1 / 3 * gmx
                    Ac[j][0][pl] -= 1 / 3 * gmn
                    Am[j][0][pl] = 0.0
                else:
                    Ap[j][0][pl] = 0.0
                    Ac[j][0][pl] += 1 / 3 * gmn
                    Am[j][0][pl] -= 1 / 3 * gmx
            elif bc.kind == "periodic":
                if bc.side < 0:
                    Am[j][0][pl] = 0.0
                else:
                    Ap[j][0][pl] = 0.0
        Ap[j][0] /= 2.0 * dxt[j]
        Ac[j][0] /= 2.0 * dxt[j]
        Am[j][0] /= 2.0 * dxt[j]


def grad_coeffs(phi: Tensor, dx: list[float], bcs: list[FaceBC]):
    """fdc.py:479-492."""
    App, Ap, Ac, Am, Amm = _defaults(phi, 1.0, 0.0, -1.0)
    _central_edit(phi, dx, bcs, Ap, Ac, Am)
    return [App, Ap, Ac, Am, Amm]


def _adv_tensor(u, phi: Tensor) -> Tensor:
    """fdc.py:775-792."""
    if isinstance(u, (float, int)):
        return torch.ones_like(phi) * float(u)
    assert u.shape == phi.shape, "adv shape must match var_i shape"
    return u


def div_coeffs(u, phi: Tensor, dx: list[float], bcs: list[FaceBC], limiter: str,
               rz_xs: list[Tensor] | None = None):
    """fdc.py:622-664, 708-772.  limiter: "none" (central) | "upwind" (reference formula)
    | "upwind_fd" (NOT in the reference: the first-order upwind difference its own test
    intends, tests/test_fdm.py:239 — parity unpinned)."""
    adv = _adv_tensor(u, phi)
    App, Ap, Ac, Am, Amm = _defaults(phi, 1.0, 0.0, -1.0)
    nd = phi.dim() - 1
    a0 = adv[0]
    if rz_xs is not None:  # tools.py:64-76: centre coefficient 2 dr / r on the r axis
        dxt0 = torch.tensor(dx, dtype=phi.dtype)
        scale = torch.nan_to_num(2 * dxt0[0] / _rz_grid(phi, rz_xs)[0], nan=0.0, posinf=0.0, neginf=0.0)
        Ac[0] = scale * torch.ones_like(phi)
    if limiter == "none":
        for j in range(nd):
            Ap[j][0] *= torch.roll(a0, -1, dims=j)
            Ac[j][0] *= a0
            Am[j][0] *= torch.roll(a0, 1, dims=j)
        if any(b.kind in ("neumann", "symmetry") for b in bcs):
            # fdc.py:741 hands a (*nx) tensor to code that indexes gamma[dim][mask]
            raise IndexError("central Div with Neumann/Symmetry faces (fdc.py:583-584)")
        _central_edit(phi, dx, bcs, Ap, Ac, Am)
    elif limiter == "upwind":
        zeros = torch.zeros_like(a0)
        for j in range(nd):
            Ap[j][0] = 2.0 * torch.min(a0, zeros)
            Ac[j][0] *= 2.0 * a0
            Am[j][0] = 2.0 * torch.max(a0, zeros)
    elif limiter == "upwind_fd":
        zeros = torch.zeros_like(a0)
        dxt = torch.tensor(dx, dtype=phi.dtype)
        up, um = torch.max(a0, zeros), torch.min(a0, zeros)
        for j in range(nd):
            Ap[j][0] = um / dxt[j]
            Ac[j][0] = (up - um) / dxt[j]
            Am[j][0] = -up / dxt[j]
    else:
        raise RuntimeError(f"{limiter=} is an unknown limiter type.")
    return [App, Ap, Ac, Am, Amm]


# ---------------------------------------------------------------------------------------
# stencil application  (fdc.py:67-118, 171-200)
# ---------------------------------------------------------------------------------------
def _axis_sum(coeffs, phi: Tensor, axis: int) -> Tensor:
    summed = torch.zeros_like(phi[0])
    for i, c in enumerate(coeffs):
        summed += c[axis][0] * torch.roll(phi[0], -2 + i, axis)
    return summed


def apply_scalar_op(coeffs, phi: Tensor) -> Tensor:
    """Laplacian (fdc.py:103-108) and Div (fdc.py:93-102) of a scalar field coincide."""
    out = torch.zeros_like(phi)
    for j in range(phi.dim() - 1):
        out[0] += _axis_sum(coeffs, phi, j)
    return out


def apply_grad(coeffs, phi: Tensor) -> Tensor:
    """fdc.py:80-87 -> (1, mesh.dim, *nx)."""
    return torch.stack(
        [torch.stack([_axis_sum(coeffs, phi, j) for j in range(phi.dim() - 1)])]
    )


# ---------------------------------------------------------------------------------------
# edge=True: one-sided differences on the domain faces (fdc.py:203-366, xyz branches)
# ---------------------------------------------------------------------------------------
def _edge_planes(nd: int, axis: int, lower: bool):
    idx = (0, 1, 2, 3) if lower else (-1, -2, -3, -4)
    return [_plane(nd, axis, i) for i in idx]


def edge_laplacian(out: Tensor, phi: Tensor, dx: list[float]) -> Tensor:
    """fdc.py:223-258.  NOTE the reference REPLACES the whole Laplacian on a face cell by the
    one-sided second derivative along that face's axis; later axes overwrite shared edges."""
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    v = phi[0]
    for a in range(nd):
        for lower in (True, False):
            p0, p1, p2, p3 = _edge_planes(nd, a, lower)
            out[0][p0] = (2.0 * v[p0] - 5.0 * v[p1] + 4.0 * v[p2] - v[p3]) / (dxt[a] ** 2)
    return out


def edge_grad(out: Tensor, phi: Tensor, dx: list[float]) -> Tensor:
    """fdc.py:260-288: component `a` on the faces normal to axis `a`."""
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    v = phi[0]
    for a in range(nd):
        p0, p1, p2, _ = _edge_planes(nd, a, True)
        out[0][a][p0] = -(3 / 2 * v[p0] - 2.0 * v[p1] + 1 / 2 * v[p2]) / (dxt[a])
        p0, p1, p2, _ = _edge_planes(nd, a, False)
        out[0][a][p0] = (3 / 2 * v[p0] - 2.0 * v[p1] + 1 / 2 * v[p2]) / (dxt[a])
    return out


def apply_div_edge(coeffs, phi: Tensor, dx: list[float], u: float) -> Tensor:
    """fdc.py:93-102 with 290-348: each axis' term is replaced on that axis' faces before the
    axes are summed (constant advection speed)."""
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    v = phi[0]
    adv = torch.ones_like(v) * u
    out = torch.zeros_like(phi)
    for a in range(nd):
        disc = _axis_sum(coeffs, phi, a)
        p0, p1, p2, _ = _edge_planes(nd, a, True)
        disc[p0] = -(3 / 2 * v[p0] - 2.0 * v[p1] + 1 / 2 * v[p2]) / (dxt[a]) * adv[p0]
        p0, p1, p2, _ = _edge_planes(nd, a, False)
        disc[p0] = (3 / 2 * v[p0] - 2.0 * v[p1] + 1 / 2 * v[p2]) / (dxt[a]) * adv[p0]
        out[0] += disc
    return out


def jacobian(phi: Tensor, dx: list[float]) -> list[Tensor]:
    """fdc.py:896-914: edge=True central gradient of a BC-less container field."""
    return list(edge_grad(apply_grad(grad_coeffs(phi, dx, []), phi), phi, dx)[0])


def hessian(phi: Tensor, dx: list[float]) -> dict[tuple[int, int], Tensor]:
    """fdc.py:917-944: gradient of every Jacobian component, upper triangle kept."""
    jac = jacobian(phi, dx)
    out = {}
    for i, j in enumerate(jac):
        hi = jacobian(j.unsqueeze(0).clone(), dx)
        for k, h in enumerate(hi):
            if i <= k:
                out[(i, k)] = h
    return out


# ---------------------------------------------------------------------------------------
# rhs adjustment (fdc.py:425-458, 505-540, 666-694)
# ---------------------------------------------------------------------------------------
def laplacian_rhs_adjust(phi, dx, bcs, rz_xs: list[Tensor] | None = None):
    out = torch.zeros_like(phi)
    nd = phi.dim() - 1
    dxt = torch.tensor(dx, dtype=phi.dtype)
    grid = _rz_grid(phi, rz_xs) if rz_xs is not None else None
    for j in range(nd):
        for bc in bcs:
            if bc.kind != "neumann":
                continue
            n = phi.shape[1 + bc.axis]
            pl = _plane(nd, bc.axis, _face_idx(bc, n, 1))
            alpha = torch.zeros_like(out[0][pl])
            if grid is not None:  # fdc.py:440-448
                dr = dxt[j] if j == 0 else 0.0
                alpha = torch.nan_to_num(1 / 3 * dr / grid[j][pl], nan=0.0, posinf=0.0, neginf=0.0)
            nvec = torch.zeros(3, dtype=phi.dtype)
            nvec[bc.axis] = bc.side
            at_bc = _bc_scalar(bc)
            if isinstance(at_bc, Tensor) and at_bc.numel() > 1:
                at_bc = at_bc.reshape(out[0][pl].shape)
            out[0][pl] += (2 / 3 - alpha) * (at_bc * nvec[j]) / dxt[j]
    return out


def _grad_like_rhs_adjust(phi, bcs, gamma_min, gamma_max):
    out = torch.zeros_like(phi)
    nd = phi.dim() - 1
    for j in range(nd):
        for bc in bcs:
            if bc.kind != "neumann":
                continue
            n = phi.shape[1 + bc.axis]
            pl = _plane(nd, bc.axis, _face_idx(bc, n, 1))
            nvec = torch.zeros(3, dtype=phi.dtype)
            nvec[bc.axis] = bc.side
            at_bc = _bc_scalar(bc)
            if isinstance(at_bc, Tensor) and at_bc.numel() > 1:
                at_bc = at_bc.reshape(out[0][pl].shape)
            g = gamma_max if bc.side < 0 else gamma_min
            out[0][pl] -= (1 / 3) * (at_bc * nvec[j]) * g[0][pl]
    return out


def grad_rhs_adjust(phi, dx, bcs):
    ones = torch.ones_like(phi)
    return _grad_like_rhs_adjust(phi, bcs, ones, ones)


def div_rhs_adjust(u, phi, dx, bcs, limiter):
    adv = _adv_tensor(u, phi)
    if limiter == "none":
        return _grad_like_rhs_adjust(phi, bcs, 2.0 * adv, 2.0 * adv)
    if limiter == "upwind":
        z = torch.zeros_like(phi)
        return _grad_like_rhs_adjust(
            phi, bcs, 2.0 * torch.min(adv, z), 2.0 * torch.max(adv, z)
        )
    if limiter == "upwind_fd":  # new scheme: Dirichlet/periodic only, no adjustment
        return torch.zeros_like(phi)
    raise RuntimeError(f"{limiter=} is an unknown limiter type.")


# ---------------------------------------------------------------------------------------
# equations:  sum_k sign_k * param_k * Op_k(phi)   (fdm.py:124-312, ops.py:47-154)
# ---------------------------------------------------------------------------------------
@dataclass
class Term:
    kind: str  # laplacian | grad | div | ddt
    sign: float = 1.0
    param: Any = None  # laplacian/grad: float|None ; div: advection (float|Tensor) ; ddt: dt
    limiter: str = "none"
    coeffs: Any = field(default=None, repr=False)


@dataclass
class Equation:
    terms: list[Term]
    dx: list[float]
    xs: list[Tensor]
    bcs: list[FaceBC]
    rz: bool = False  # axisymmetric (r, z) mesh: Cylinder geometry
    # nonlinear advection, Term("div", param="self") == fdm.div(var, var): the Div coefficients are
    # rebuilt from the LIVE unknown on every application (fdm.py:306-312); the solvers below move
    # `live` where the reference rebinds var (linalg.py:122-125, 236-238, 253-256)
    live: Any = field(default=None, repr=False)

    def set_live(self, x: Tensor) -> None:
        self.live = x

    def build(self, phi: Tensor) -> "Equation":
        rz_xs = self.xs if self.rz else None
        for t in self.terms:
            if t.kind == "laplacian":
                t.coeffs = laplacian_coeffs(phi, self.dx, self.bcs, rz_xs)
            elif t.kind == "grad":
                t.coeffs = grad_coeffs(phi, self.dx, self.bcs)
            elif t.kind == "div":
                if isinstance(t.param, str):  # "self": built per application in aop()
                    self.live = phi
                    t.coeffs = None
                else:
                    t.coeffs = div_coeffs(t.param, phi, self.dx, self.bcs, t.limiter, rz_xs)
            elif t.kind == "ddt":
                # implicit Euler: the linear part of (phi - phi_old)/dt, as a multiplication by
                # 1/dt rounded in the field dtype (NOT in the reference, see implicit_euler_step)
                t.coeffs = torch.ones(1, dtype=phi.dtype) / t.param
            else:
                raise ValueError(t.kind)
        return self

    def adjust_rhs(self, phi: Tensor, rhs: Tensor) -> Tensor:
        """ops.py:63-77 — in place on the caller's tensor, every term contributes."""
        for t in self.terms:
            if t.kind == "ddt":  # ops.py:70-71 skips Ddt
                continue
            if t.kind == "laplacian":
                rhs += laplacian_rhs_adjust(phi, self.dx, self.bcs, self.xs if self.rz else None)
            elif t.kind == "grad":
                rhs += grad_rhs_adjust(phi, self.dx, self.bcs)
            else:
                u = phi if isinstance(t.param, str) else t.param
                rhs += div_rhs_adjust(u, phi, self.dx, self.bcs, t.limiter)
        return rhs

    def aop(self, phi: Tensor) -> Tensor:
        """ops.py:122-154."""
        res = torch.zeros_like(phi)
        for t in self.terms:
            if t.kind == "ddt":  # skipped in the loop, added after it (ops.py:133-134,151-152)
                continue
            if t.kind == "grad":
                ax = apply_grad(t.coeffs, phi)
                if t.param is not None:
                    ax = ax * t.param
                ax = (ax * t.sign).view(phi.size())  # only legal in 1-D (ops.py:145-147)
            else:
                coeffs = t.coeffs
                if t.kind == "div" and isinstance(t.param, str):
                    coeffs = div_coeffs(self.live, phi, self.dx, self.bcs, t.limiter, self.xs if self.rz else None)
                ax = apply_scalar_op(coeffs, phi)
                if t.kind == "laplacian" and t.param is not None:
                    ax = ax * t.param
                ax = ax * t.sign
            res += ax
        for t in self.terms:
            if t.kind == "ddt":
                res += t.coeffs * phi
        return res

    def diag(self, phi: Tensor) -> Tensor:
        """Centre coefficient of the summed operator (for Jacobi; not in the reference)."""
        res = torch.zeros_like(phi)
        for t in self.terms:
            if t.kind == "ddt":
                continue
            d = torch.zeros_like(phi)
            for j in range(phi.dim() - 1):
                d[0] += t.coeffs[2][j][0]
            if t.kind in ("laplacian", "grad") and t.param is not None:
                d = d * t.param
            res += d * t.sign
        for t in self.terms:
            if t.kind == "ddt":
                res += t.coeffs * torch.ones_like(phi)
        return res


# ---------------------------------------------------------------------------------------
# solvers  (linalg.py)
# ---------------------------------------------------------------------------------------
def _nan_to_num(t: Tensor) -> Tensor:  # linalg.py:302-305
    return torch.nan_to_num(t, nan=0.0, posinf=0.0, neginf=0.0)


def tolerance_check(a: Tensor, b: Tensor) -> float:  # linalg.py:321-338
    tol = torch.zeros(a.shape[0], dtype=a.dtype)
    for d in range(a.shape[0]):
        tol[d] = torch.linalg.norm(a[d] - b[d])
    if torch.isnan(tol) or torch.isinf(tol):
        raise RuntimeError(f"Invalid tolerance detected! tol: {tol}")
    return torch.max(tol).item()


def cg(eq: Equation, x: Tensor, rhs: Tensor, tolerance: float, max_it: int):
    """linalg.py:74-159.  Returns (x, report, x_prev)."""
    axes = list(range(1, x.dim()))
    sl = solver_region(x.dim() - 1, eq.bcs)
    tol, itr = 1.0, 0
    apply_bcs(x, eq.xs, eq.bcs)
    eq.set_live(x)
    Ad = torch.zeros_like(rhs)
    r = torch.zeros_like(x)
    r[0][sl] = rhs[0][sl] - eq.aop(x)[0][sl]
    d = r.clone()
    x_old = x
    while tol > tolerance:
        x_old = x.clone()
        Ad[0][sl] = eq.aop(d)[0][sl]
        alpha = _nan_to_num(torch.sum(r * r, dim=axes) / torch.sum(d * Ad, dim=axes))
        x = x + alpha * d
        apply_bcs(x, eq.xs, eq.bcs)
        eq.set_live(x)
        beta_denom = torch.sum(r * r, dim=axes)
        r -= alpha * Ad
        tol = tolerance_check(x, x_old)
        beta = torch.sum(r * r, dim=axes) / beta_denom
        d = r + beta * d
        itr += 1
        if itr > max_it:
            warnings.warn(f"Maximum iteration reached! max_it: {max_it}", RuntimeWarning)
            break
    return x, {"itr": itr, "tol": tol, "converge": itr < max_it}, x_old


def bicgstab(eq: Equation, x: Tensor, rhs: Tensor, tolerance: float, max_it: int):
    """linalg.py:162-279."""
    axes = list(range(1, x.dim()))
    sl = solver_region(x.dim() - 1, eq.bcs)
    itr = 0
    apply_bcs(x, eq.xs, eq.bcs)
    eq.set_live(x)
    r0 = torch.zeros_like(x)
    r0[0][sl] = rhs[0][sl] - eq.aop(x)[0][sl]
    r = r0.clone()
    t = torch.zeros_like(x)
    v = torch.zeros_like(x)
    p = torch.zeros_like(x)
    rho: Any = 1.0
    alpha: Any = 1.0
    omega: Any = 1.0
    rho_next = torch.sum(r0 * r0, dim=axes)
    tol = torch.sqrt(rho_next.max()).item()
    finished = False
    x_old = x
    while not finished:
        x_old = x.clone()
        beta = rho_next / rho * alpha / omega
        rho = rho_next
        p = r + beta * (p - omega * v)
        v[0][sl] = eq.aop(p)[0][sl]
        itr += 1
        alpha = _nan_to_num(rho / torch.sum(r0 * v, dim=axes))
        s = r - alpha * v
        tol = tolerance_check(r, alpha * v)
        if tol <= tolerance:
            x = x + alpha * p
            apply_bcs(x, eq.xs, eq.bcs)
            eq.set_live(x)
            finished = True
            continue
        t[0][sl] = eq.aop(s)[0][sl]
        omega = _nan_to_num(torch.sum(t * s, dim=axes) / torch.sum(t * t, dim=axes))
        rho_next = -omega * torch.sum(r0 * t, dim=axes)
        x = x + alpha * p + s * omega
        apply_bcs(x, eq.xs, eq.bcs)
        eq.set_live(x)
        r = s - omega * t
        tol = tolerance_check(s, omega * t)
        if tol <= tolerance:
            finished = True
        if itr >= max_it:
            warnings.warn(f"Maximum iteration reached! max_it: {max_it}", RuntimeWarning)
            break
    return x, {"itr": itr, "tol": tol, "converge": itr < max_it}, x_old


def jacobi(eq: Equation, x: Tensor, rhs: Tensor, tolerance: float, max_it: int):
    """NOT in the reference (linalg.py:62-69 dispatches cg/bicgstab only).  Defined from
    reference primitives: loop/exit/itr conventions of `cg`."""
    sl = solver_region(x.dim() - 1, eq.bcs)
    tol, itr = 1.0, 0
    apply_bcs(x, eq.xs, eq.bcs)
    diag = eq.diag(x)
    x_old = x
    while tol > tolerance:
        x_old = x.clone()
        eq.set_live(x)
        ax = eq.aop(x)
        x = x.clone()
        x[0][sl] = x_old[0][sl] + (rhs[0][sl] - ax[0][sl]) / diag[0][sl]
        apply_bcs(x, eq.xs, eq.bcs)
        tol = tolerance_check(x, x_old)
        itr += 1
        if itr > max_it:
            warnings.warn(f"Maximum iteration reached! max_it: {max_it}", RuntimeWarning)
            break
    return x, {"itr": itr, "tol": tol, "converge": itr < max_it}, x_old


def euler_step(eq: Equation, x: Tensor, rhs: Tensor | None, dt: float) -> Tensor:
    """NOT in the reference (fdm.Ddt is a stub, fdm.py:315-353).  Explicit Euler
    (SURVEY.md §8a A16): phi_new[sl] = phi + dt*(rhs - Aop(phi)), then BCs."""
    sl = solver_region(x.dim() - 1, eq.bcs)
    eq.set_live(x)
    ax = eq.aop(x)
    src = torch.zeros_like(x) if rhs is None else rhs
    new = x.clone()
    new[0][sl] = x[0][sl] + dt * (src[0][sl] - ax[0][sl])
    apply_bcs(new, eq.xs, eq.bcs)
    return new


def implicit_euler_step(eq: Equation, x: Tensor, rhs: Tensor | None, dt: float, method: str, tolerance: float,
                        max_it: int):
    """NOT in the reference: `fdm.Ddt` registers nothing (fdm.py:322-339); the semantics its
    failing test intends is Aop = (phi - phi_old)/dt + spatial operators (tests/test_fdm.py:275-299).
    Implicit Euler: with c = 1/dt rounded in the field dtype, solve
        c*phi_new + A_spatial(phi_new) = rhs + c*phi_old
    with the named solver, initial guess phi_old.  `eq` must hold a Term("ddt", param=dt); its
    contribution c*phi is added after the spatial operators (ops.py:151-152).  Parity unpinned."""
    c = torch.ones(1, dtype=x.dtype) / dt
    src = torch.zeros_like(x) if rhs is None else rhs
    rhs_eff = src + c * x
    solver = {"cg": cg, "bicgstab": bicgstab, "jacobi": jacobi}[method]
    return solver(eq, x.clone(), rhs_eff, tolerance, max_it)
