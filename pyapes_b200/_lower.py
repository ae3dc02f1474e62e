"""Lowering of the Python objects (Mesh / Field / BC / operator dicts) to the plain C structs of
include/pyapes_b200.h.  Host-side setup code only; nothing here runs per iteration.

Rounding contract: every coefficient the kernels multiply with is produced HERE with the same
sequence of torch operations, in the field dtype, that the reference uses to fill its
coefficient tensors (cited per function), so the device arithmetic reproduces the reference's
values bit for bit.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import torch
from torch import Tensor

from pyapes_b200 import _native as N


_KERNEL_AXES = {1: (2,), 2: (0, 2), 3: (0, 1, 2)}


def kernel_axis(mesh_axis: int, ndim: int) -> int:
    """Mesh axis -> kernel axis.  Kernel axis 2 is the contiguous one; 2-D meshes use kernel axes
    (0, 2) so that the kernels march along the first mesh axis."""
    return _KERNEL_AXES[ndim][mesh_axis]


def lower_grid(nx, bcs, slab=None) -> N.Grid:
    """Grid block + the solver region of mesh/tools.py:7-20.  `slab` (pyapes_b200.parallel)
    places a local block with ghost planes inside the global grid along kernel axis 0."""
    nd = len(nx)
    g = N.Grid()
    n, lo, hi = [1, 1, 1], [0, 0, 0], [1, 1, 1]
    for j, v in enumerate(nx):
        a = kernel_axis(j, nd)
        n[a], lo[a], hi[a] = int(v), 1, int(v) - 1
    for bc in bcs or []:
        if bc.bc_type == "periodic":
            a = kernel_axis(bc.bc_face_dim, nd)
            if bc.bc_n_dir < 0:
                lo[a] = 0
            else:
                hi[a] = n[a]
    for a in range(3):
        g.n[a], g.lo[a], g.hi[a] = n[a], lo[a], max(hi[a], lo[a])
    g.gn0, g.goff0, g.olo0, g.ohi0 = n[0], 0, 0, n[0]
    g.ndim = nd
    if slab is not None:
        assert nd == 3, "slab decomposition is along axis 0 of a 3-D mesh"
        g.gn0, g.goff0, g.olo0, g.ohi0 = slab["gn0"], slab["goff0"], slab["olo0"], slab["ohi0"]
        per0 = (lo[0] == 0, hi[0] == n[0])
        if slab["world"] > 1 and (any(per0) != bool(slab.get("periodic", False)) or per0[0] != per0[1]):
            raise ValueError(
                "pyapes_b200: Periodic BCs on the x faces of a slab-decomposed mesh need both faces periodic "
                "and SlabMesh(..., periodic=True) (wrap-around ghost planes); and vice versa"
            )
        # region along axis 0: owned planes that are inside the global slicer
        glo, ghi = (0 if lo[0] == 0 else 1), (slab["gn0"] if hi[0] == n[0] else slab["gn0"] - 1)
        g.lo[0] = max(slab["olo0"], glo - slab["goff0"])
        g.hi[0] = max(g.lo[0], min(slab["ohi0"], ghi - slab["goff0"]))
    return g


def _face_points(bc, grid) -> tuple[Tensor, Tensor]:
    """Coordinate of the face plane and of the plane one cell inside, as 0-d tensors
    (bcs.py:228-231 takes them from the meshgrid through the masks)."""
    a = bc.bc_face_dim
    g = grid[a]
    n = g.shape[a]
    pf = 0 if bc.bc_n_dir < 0 else n - 1
    p1 = (pf - bc.bc_n_dir) % n
    idx_f = tuple(pf if i == a else 0 for i in range(g.dim()))
    idx_1 = tuple(p1 if i == a else 0 for i in range(g.dim()))
    return g[idx_f], g[idx_1]


def _frozen_callable_check(bc, grid, var: Tensor, out) -> None:
    """The solvers evaluate a callable `bc_val` ONCE per solve and keep the face values fixed; the reference
    calls it at every BC application with the CURRENT iterate (bcs.py:203-206, 240-241; linalg.py:295-297).
    The two agree unless the callable reads `var`.  Probe: evaluate it again on a shifted copy of the field;
    a different answer means it depends on the iterate, which the device-side iteration cannot honour
    (SURVEY.md 8b fallback rule: raise, never run a silently different computation)."""
    probe = bc.bc_val(grid, bc.bc_mask, var * 0.5 + 1.0, bc.bc_val_opt)
    same = (torch.equal(torch.as_tensor(out), torch.as_tensor(probe)) if isinstance(out, Tensor) or isinstance(probe, Tensor)
            else out == probe)
    if not same:
        raise NotImplementedError(
            f"pyapes_b200: the callable bc_val of face {bc.bc_face!r} depends on the field itself; inside a solver "
            "the reference re-evaluates it on every iterate, which the device-resident iteration does not do. "
            "Use a callable of (grid, mask) only, or apply such a BC explicitly with BC.apply between solves."
        )


def face_value(bc, grid, var: Tensor, var_dim: int, frozen: bool = False) -> tuple[float | None, Tensor | None]:
    """(scalar, per-cell array) form of bc.bc_val (bcs.py:203-213, 240-249).  `frozen`: the caller keeps
    the values for a whole solve (see _frozen_callable_check)."""
    v = bc.bc_val
    if callable(v):
        out = v(grid, bc.bc_mask, var, bc.bc_val_opt)
        if frozen:
            _frozen_callable_check(bc, grid, var, out)
        if not isinstance(out, Tensor):
            return float(out), None
        return (float(out), None) if out.numel() == 1 else (None, out)
    if isinstance(v, list):
        return float(v[var_dim]), None
    if isinstance(v, (int, float)):
        return float(v), None
    if isinstance(v, Tensor):
        return (float(v), None) if v.numel() == 1 else (None, v)
    return None, None


def lower_face(bc, grid, var: Tensor, var_dim: int, nd: int, frozen: bool = False) -> tuple[N.FaceBC, Any]:
    """One pa_face_bc.  Returns the struct and the tensor (if any) that must stay alive."""
    f = N.FaceBC()
    f.axis = kernel_axis(bc.bc_face_dim, nd)
    f.side = bc.bc_n_dir
    f.kind = N.BC_KIND[bc.bc_type]
    f.value = 0.0
    f.values = None
    keep = None
    dt, dev = var.dtype, var.device
    if bc.bc_type == "dirichlet":
        assert bc.bc_val is not None, "BC: bc_val is not specified!"
        s, arr = face_value(bc, grid, var, var_dim, frozen)
        if arr is None and s is None:
            raise TypeError("Dirichlet: bc_val must be float, int, callable or list!")
        if arr is not None:
            keep = arr.to(device=dev, dtype=dt).contiguous().reshape(-1)
        else:
            f.value = s
    elif bc.bc_type == "neumann":
        assert bc.bc_val is not None, "BC: bc_val is not specified!"
        s, arr = face_value(bc, grid, var, var_dim, frozen)
        if arr is None and s is None:
            raise TypeError("Neumann: bc_val must be float, int, callable or list!")
        xf, x1 = _face_points(bc, grid)
        dxv = xf - x1  # grid[mask] - grid[mask_prev]                          bcs.py:228-231
        if arr is not None:
            c = 2 / 3 * arr.to(device=dev, dtype=dt) * dxv * bc.bc_n_dir  # bcs.py:252
            keep = c.contiguous().reshape(-1)
        else:
            f.value = float(2 / 3 * s * dxv * bc.bc_n_dir)
    if keep is not None:
        N.require_cuda(keep, "boundary value array")
        f.values = keep.data_ptr()
    return f, keep


def lower_faces(bcs, grid, var: Tensor, var_dim: int, nd: int, frozen: bool = False):
    """`frozen=True` from the solvers: the face values stay fixed for the whole solve."""
    arr = (N.FaceBC * max(len(bcs), 1))()
    keep = []
    for i, bc in enumerate(bcs):
        arr[i], k = lower_face(bc, grid, var, var_dim, nd, frozen)
        keep.append(k)
    return arr, len(bcs), keep


# ---------------------------------------------------------------------------------------
# coefficient classes of the constant-coefficient stars
# ---------------------------------------------------------------------------------------
class StarCoeffs:
    """What the reference stores as 15 full-size tensors (`[App, Ap, Ac, Am, Amm]`, each a
    list over mesh axes, tools.py:29-112) collapses, for constant coefficients, to three
    numbers per axis and per plane class (index 1 / index n-2 / elsewhere).  `vec[j]` keeps the
    per-axis 1-D coefficient vectors [Ap, Ac, Am] (length n_j) they were read from."""

    kind = N.OP_STAR

    def __init__(self, vecs: list[list[Tensor]], name: str, table_axes=()):
        self.vec = vecs
        self.name = name
        self.adv = None
        # axes whose coefficients vary with the index (rz: the r axis) go to the kernel as a
        # per-index table instead of three classes
        self.table_axes = tuple(table_axes)

    def table(self, j: int) -> Tensor:
        ap, ac, am = self.vec[j]
        return torch.stack([ap, ac, am], dim=1).contiguous()

    def classes(self, j: int) -> list[list[float]]:
        ap, ac, am = self.vec[j]
        n = ap.numel()
        pick = (0, 1 % n, (n - 2) % n)
        return [[float(ap[i]), float(ac[i]), float(am[i])] for i in pick]


class FieldCoeffs:
    """Div with a tensor / Field advection speed: coefficients are formed in-kernel from the
    live field (fdc.py:708-772); only scalars and periodic flags are kept here."""

    def __init__(self, kind: int, adv: Tensor, dx: list[float], zero_am_lo, zero_ap_hi, name: str):
        self.kind, self.adv, self.dx = kind, adv, dx
        self.zero_am_lo, self.zero_ap_hi = zero_am_lo, zero_ap_hi
        self.name = name


def _axis_bcs(bcs, j):
    return [bc for bc in (bcs or []) if bc.bc_face_dim == j]


def laplacian_star(nx, dx: list[float], bcs, dtype, rz_x=None) -> StarCoeffs:
    """fdc.py:375-423 with tools.py:79-85 on one 1-D vector per axis.  `rz_x` = the (r, z) node
    coordinate vectors of an axisymmetric mesh: r-axis coefficients 1 +- dr/(2r) (tools.py:86-108)
    and the Neumann/Symmetry edits with alpha = 2/3 dr/r (fdc.py:395-403)."""
    dxt = torch.tensor(dx, dtype=dtype)
    vecs = []
    for j, n in enumerate(nx):
        ap, ac, am = torch.ones(n, dtype=dtype), -2.0 * torch.ones(n, dtype=dtype), torch.ones(n, dtype=dtype)
        if rz_x is not None and j == 0:
            scale = torch.nan_to_num(dxt[0] / (2 * rz_x[0]), nan=0.0, posinf=0.0, neginf=0.0)
            ap, am = (1 + scale) * ap, (1 - scale) * am
        for bc in _axis_bcs(bcs, j):
            if bc.bc_type in ("neumann", "symmetry"):
                alpha = torch.zeros(1, dtype=dtype)
                if rz_x is not None:
                    dr = dxt[j] if j == 0 else 0.0
                    r = rz_x[j][1 % n] if bc.bc_n_dir < 0 else rz_x[j][(n - 2) % n]
                    alpha = torch.nan_to_num(2 / 3 * dr / r, nan=0.0, posinf=0.0, neginf=0.0).reshape(1)
                if bc.bc_n_dir < 0:
                    ap[1 % n] = 2 / 3 + alpha
                    ac[1 % n] = -(2 / 3 + alpha)
                    am[1 % n] = 0.0
                else:
                    ap[(n - 2) % n] = 0.0
                    ac[(n - 2) % n] = -(2 / 3 + alpha)
                    am[(n - 2) % n] = 2 / 3 + alpha
        ap /= dxt[j] ** 2
        ac /= dxt[j] ** 2
        am /= dxt[j] ** 2
        vecs.append([ap, ac, am])
    return StarCoeffs(vecs, "Laplacian", table_axes=(0,) if rz_x is not None else ())


def _central_star(nx, dx, bcs, dtype, ap0, ac0, am0, gamma, name, allow_ns=True) -> StarCoeffs:
    """fdc.py:543-609: central-difference edits with a constant gamma, then / (2 dx)."""
    dxt = torch.tensor(dx, dtype=dtype)
    vecs = []
    for j, n in enumerate(nx):
        ap, ac, am = ap0(n), ac0(n), am0(n)
        g = torch.ones(1, dtype=dtype) * gamma
        for bc in _axis_bcs(bcs, j):
            if bc.bc_type in ("neumann", "symmetry"):
                if not allow_ns:
                    raise IndexError(
                        "FDC Div: central scheme with Neumann/Symmetry faces is not usable "
                        "(the reference raises IndexError in fdc.py:583-584 as well)"
                    )
                if bc.bc_n_dir < 0:
                    ap[1 % n] += (1 / 3 * g)[0]
                    ac[1 % n] -= (1 / 3 * g)[0]
                    am[1 % n] = 0.0
                else:
                    ap[(n - 2) % n] = 0.0
                    ac[(n - 2) % n] += (1 / 3 * g)[0]
                    am[(n - 2) % n] -= (1 / 3 * g)[0]
            elif bc.bc_type == "periodic":
                if bc.bc_n_dir < 0:
                    am[1 % n] = 0.0
                else:
                    ap[(n - 2) % n] = 0.0
        ap /= 2.0 * dxt[j]
        ac /= 2.0 * dxt[j]
        am /= 2.0 * dxt[j]
        vecs.append([ap, ac, am])
    return StarCoeffs(vecs, name)


def grad_star(nx, dx, bcs, dtype) -> StarCoeffs:
    """fdc.py:479-492."""
    one = lambda n: torch.ones(n, dtype=dtype)  # noqa: E731
    return _central_star(nx, dx, bcs, dtype, one, lambda n: torch.zeros(n, dtype=dtype),
                         lambda n: -1.0 * one(n), 1.0, "Grad")


def div_star_const(u: float, nx, dx, bcs, dtype, limiter: str, rz_x=None) -> StarCoeffs:
    """Div with a constant advection speed (fdc.py:622-664, 708-772).  rz: the centre coefficient
    on the r axis starts as 2 dr / r (tools.py:64-76)."""
    adv = torch.ones(1, dtype=dtype) * u  # fdc.py:779
    if rz_x is not None:
        return _div_star_const_rz(u, adv, nx, dx, bcs, dtype, limiter, rz_x)
    if limiter == "none":
        if any(bc.bc_type in ("neumann", "symmetry") for bc in (bcs or [])):
            raise IndexError(
                "FDC Div: central scheme with Neumann/Symmetry faces is not usable "
                "(the reference raises IndexError in fdc.py:583-584 as well)"
            )
        a = adv[0]
        return _central_star(nx, dx, bcs, dtype,
                             lambda n: torch.ones(n, dtype=dtype) * a,
                             lambda n: torch.zeros(n, dtype=dtype) * a,
                             lambda n: -1.0 * torch.ones(n, dtype=dtype) * a, u, "Div", allow_ns=False)
    vecs = []
    zero = torch.zeros(1, dtype=dtype)
    if limiter == "upwind":  # fdc.py:765-770: no 1/dx, centre coefficient 0 * 2u
        for n in nx:
            ap = (2.0 * torch.min(adv, zero)).expand(n).clone()
            ac = (torch.zeros(1, dtype=dtype) * (2.0 * adv)).expand(n).clone()
            am = (2.0 * torch.max(adv, zero)).expand(n).clone()
            vecs.append([ap, ac, am])
    elif limiter == "upwind_fd":  # not in the reference: u+ (phi - phi[-1])/dx + u- (phi[+1] - phi)/dx
        dxt = torch.tensor(dx, dtype=dtype)
        up, um = torch.max(adv, zero), torch.min(adv, zero)
        for j, n in enumerate(nx):
            vecs.append([(um / dxt[j]).expand(n).clone(), ((up - um) / dxt[j]).expand(n).clone(),
                         (-up / dxt[j]).expand(n).clone()])
    elif limiter == "quick":
        raise NotImplementedError("FDC Div: quick scheme is not implemented yet.")
    else:
        raise RuntimeError(f"FDC Div: {limiter=} is an unknown limiter type.")
    return StarCoeffs(vecs, "Div")


def _div_star_const_rz(u, adv, nx, dx, bcs, dtype, limiter, rz_x) -> StarCoeffs:
    dxt = torch.tensor(dx, dtype=dtype)
    zero = torch.zeros(1, dtype=dtype)
    vecs = []
    for j, n in enumerate(nx):
        ap, ac, am = torch.ones(n, dtype=dtype), torch.zeros(n, dtype=dtype), -1.0 * torch.ones(n, dtype=dtype)
        if j == 0:
            ac = torch.nan_to_num(2 * dxt[0] / rz_x[0], nan=0.0, posinf=0.0, neginf=0.0) * torch.ones(n, dtype=dtype)
        if limiter == "none":
            if any(bc.bc_type in ("neumann", "symmetry") for bc in (bcs or [])):
                raise IndexError("FDC Div: central scheme with Neumann/Symmetry faces is not usable "
                                 "(the reference raises IndexError in fdc.py:583-584 as well)")
            ap, ac, am = ap * adv[0], ac * adv[0], am * adv[0]
            for bc in _axis_bcs(bcs, j):
                if bc.bc_type == "periodic":
                    if bc.bc_n_dir < 0:
                        am[1 % n] = 0.0
                    else:
                        ap[(n - 2) % n] = 0.0
            ap, ac, am = ap / (2.0 * dxt[j]), ac / (2.0 * dxt[j]), am / (2.0 * dxt[j])
        elif limiter == "upwind":
            ap = (2.0 * torch.min(adv, zero)).expand(n).clone()
            ac = ac * (2.0 * adv)
            am = (2.0 * torch.max(adv, zero)).expand(n).clone()
        else:
            raise NotImplementedError(f"pyapes_b200: limiter {limiter!r} on rz meshes")
        vecs.append([ap, ac, am])
    return StarCoeffs(vecs, "Div", table_axes=(0,))


def div_field(adv: Tensor, nx, dx, bcs, limiter: str) -> FieldCoeffs:
    nd = len(nx)
    zl, zh = [0] * 3, [0] * 3
    if limiter == "none":
        if any(bc.bc_type in ("neumann", "symmetry") for bc in (bcs or [])):
            raise IndexError(
                "FDC Div: central scheme with Neumann/Symmetry faces is not usable "
                "(the reference raises IndexError in fdc.py:583-584 as well)"
            )
        for bc in bcs or []:
            if bc.bc_type == "periodic":
                a = kernel_axis(bc.bc_face_dim, nd)
                if bc.bc_n_dir < 0:
                    zl[a] = 1
                else:
                    zh[a] = 1
        kind = N.OP_DIV_CENTRAL_FIELD
    elif limiter == "upwind":
        kind = N.OP_DIV_UPWIND_FIELD
    elif limiter == "upwind_fd":
        kind = N.OP_DIV_UPWINDFD_FIELD
    elif limiter == "quick":
        raise NotImplementedError("FDC Div: quick scheme is not implemented yet.")
    else:
        raise RuntimeError(f"FDC Div: {limiter=} is an unknown limiter type.")
    return FieldCoeffs(kind, adv, dx, zl, zh, "Div")


def lower_op(coeffs, nd: int, dtype, sign: float = 1.0, param=None, field_shape=None, edge: int = 0,
             dx=None, adv_const: float = 0.0, field_device=None) -> tuple[N.Op, Any]:
    """StarCoeffs / FieldCoeffs -> pa_op.  `param` is None, a float, or a Tensor broadcastable to
    the field (the reference multiplies the stencil result by it elementwise, fdm.py:169)."""
    op = N.Op()
    op.kind = coeffs.kind
    op.sign = float(sign)
    keep = None
    pkeep = None
    if isinstance(param, Tensor):
        if param.numel() == 1:
            param = float(param)
        else:
            assert field_shape is not None
            pkeep = torch.broadcast_to(param, field_shape).to(dtype).contiguous()
            N.require_cuda(pkeep, "operator coefficient tensor")
            op.param_field = pkeep.data_ptr()
            param = 1.0
    op.has_param = 0 if param is None else 1
    op.param = 1.0 if param is None else float(param)
    op.edge = int(edge)
    op.adv_const = float(torch.ones(1, dtype=dtype)[0] * adv_const)
    if dx is not None:
        dxt0 = torch.tensor(dx, dtype=dtype)
        for j in range(nd):
            op.dx[kernel_axis(j, nd)] = float(dxt0[j])
    tkeep = []
    if isinstance(coeffs, StarCoeffs):
        for j in range(nd):
            a = kernel_axis(j, nd)
            cls = coeffs.classes(j)
            for c in range(3):
                for k in range(3):
                    op.coef[a][c][k] = cls[c][k]
            if j in coeffs.table_axes:
                dev = field_device if field_device is not None else "cuda"
                tab = coeffs.table(j).to(device=dev, dtype=dtype)
                op.coef_tab[a] = tab.data_ptr()
                tkeep.append(tab)
    else:
        adv = coeffs.adv
        N.require_cuda(adv, "advection field")
        keep = adv
        op.adv = adv.data_ptr()
        dxt = torch.tensor(coeffs.dx, dtype=dtype)
        for j in range(nd):
            a = kernel_axis(j, nd)
            op.two_dx[a] = float(2.0 * dxt[j])
            op.dx[a] = float(dxt[j])
        for a in range(3):
            op.zero_am_lo[a] = coeffs.zero_am_lo[a]
            op.zero_ap_hi[a] = coeffs.zero_ap_hi[a]
    return op, (keep, pkeep, tkeep)
