"""Explicit finite-difference discretisers `FDC().laplacian / .grad / .div`
(reference: pyapes/solver/fdc.py).

Public surface kept: `Discretizer.build_A_coeffs / adjust_rhs / apply / __call__ / reset /
set_config`, the attributes `A_coeffs`, `rhs_adj`, and the class-level singletons on `FDC`.

Underneath, `apply` is one CUDA launch (csrc: k_apply / k_grad) with the boundary-adjacent
coefficient edits resolved in-kernel, instead of 5 rolls x 5 multiplies x 5 adds per axis
against full-size coefficient tensors (fdc.py:171-200).  `edge=True`, `DiffFlux`, `jacobian`
and `hessian` are "next" rows of SURVEY.md §8(f) and raise NotImplementedError.
"""
from __future__ import annotations

import warnings
from typing import Any

import torch
from torch import Tensor

from pyapes_b200 import _lower as L
from pyapes_b200 import _native as N
from pyapes_b200.solver.types import DiscretizerConfigType, DivConfigType
from pyapes_b200.variables import Field


def _rz_x(var: Field):
    """Host node-coordinate vectors of an axisymmetric mesh (None for Cartesian ones)."""
    return var.mesh._x_host if var.mesh.coord_sys == "rz" else None


def _scalar_field(var: Field, who: str) -> None:
    if var.dim != 1:
        raise NotImplementedError(
            f"pyapes_b200 {who}: only scalar fields (Field.dim == 1) are on the CUDA path; the "
            "reference's solvers do not work for vector fields either (SURVEY.md §0 item 5)."
        )


def _plane(var: Field, bc, shift: int):
    a, n = bc.bc_face_dim, var.nx[bc.bc_face_dim]
    base = 0 if bc.bc_n_dir < 0 else n - 1
    idx: list[Any] = [slice(None)] * var.mesh.dim
    idx[a] = (base - bc.bc_n_dir * shift) % n
    return tuple(idx)


def _return_bc_val(bc, var: Field, dim: int, shape) -> Tensor | float:
    """fdc.py:803-817."""
    v = bc.bc_val
    if callable(v):
        out = v(var.mesh.grid, bc.bc_mask, var(), bc.bc_n_vec)
        return out.reshape(shape) if isinstance(out, Tensor) and out.numel() > 1 else out
    if isinstance(v, list):
        return v[dim]
    if isinstance(v, (float, int)):
        return v
    if v is None:
        return 0.0
    raise ValueError(f"Unknown boundary condition value: {v}")


def _owns_face(var: Field, bc) -> bool:
    """Slab-decomposed meshes: the two faces normal to axis 0 exist on the edge ranks only."""
    slab = getattr(var.mesh, "slab", None)
    if slab is None or bc.bc_face_dim != 0:
        return True
    return slab["rank"] == (0 if bc.bc_n_dir < 0 else slab["world"] - 1)


def _grad_like_rhs(var: Field, gamma_min: Tensor, gamma_max: Tensor) -> Tensor:
    """fdc.py:505-540: Neumann faces only; lower faces take gamma_max, upper gamma_min."""
    out = torch.zeros_like(var())
    if var.bcs is None:
        return out
    for j in range(var.mesh.dim):
        for bc in var.bcs:
            if bc.bc_type != "neumann" or not _owns_face(var, bc):
                continue
            pl = _plane(var, bc, 1)
            at_bc = _return_bc_val(bc, var, 0, out[0][pl].shape)
            g = gamma_max if bc.bc_n_dir < 0 else gamma_min
            out[0][pl] -= (1 / 3) * (at_bc * bc.bc_n_vec[j]) * g[0][pl]
    return out


class Discretizer:
    A_coeffs: Any = None
    rhs_adj: Tensor | None = None
    _op_type: str = "Discretizer"
    _config: DiscretizerConfigType | None = None

    op_type = property(lambda self: self._op_type)
    config = property(lambda self: self._config)

    def _edge(self) -> bool:
        if self.config is not None and self.op_type.lower() in self.config:
            return bool(self.config[self.op_type.lower()]["edge"])  # type: ignore[literal-required]
        if self.config is None:
            warnings.warn("FDC: config is not defined! Using default config (edge=False).")
        return False

    def apply(self, A_coeffs, var: Field) -> Tensor:
        """The stencil (fdc.py:67-118) as one kernel launch."""
        assert A_coeffs is not None, "FDC: A_A_coeffs is not defined!"
        edge = self._edge()
        _scalar_field(var, f"FDC.{self.op_type}")
        edge_code, adv_const = 0, 0.0
        if edge:
            # one-sided differences on the domain faces (fdc.py:203-366)
            need = 4 if self.op_type == "Laplacian" else 3
            if min(var.nx) < need:
                raise IndexError(f"FDC {self.op_type}: edge=True needs at least {need} points per axis")
            if self.op_type == "Div":
                if var.mesh.dim > 1:
                    # the reference indexes var[dim] on a scalar field here (fdc.py:303-307)
                    raise IndexError("index 1 is out of bounds for dimension 0 with size 1")
                adv = getattr(self, "var_addition", None)
                if adv is None:
                    adv = 1.0
                if not isinstance(adv, (float, int)):
                    raise NotImplementedError("pyapes_b200: edge=True Div needs a constant advection speed")
                edge_code, adv_const = 2, float(adv)
            else:
                edge_code = 1
        phi = var()
        N.require_cuda(phi, "field")
        nd = var.mesh.dim
        slab = getattr(var.mesh, "slab", None)
        if slab is not None and slab["world"] > 1:
            # Slab-decomposed mesh: refresh the ghost planes, then apply on the local block.  Every OWNED cell
            # gets the single-GPU value, with one exception: edge=False on a NON-periodic slab axis leaves the two
            # global x-boundary planes with a local instead of the global wrap-around neighbour -- values that are
            # torch.roll artefacts in the reference as well (it never uses them); edge=True replaces them by the
            # one-sided formulas and is identical everywhere.  Ghost planes of the result are not meaningful.
            from pyapes_b200 import parallel

            gl = L.lower_grid(var.nx, var.bcs, slab)
            N.check(N.lib().pa_halo_exchange(gl, N.dtype_code(phi.dtype), phi.data_ptr(), 1 if slab["periodic"] else 0,
                                             parallel.get_comm(phi.device), slab["rank"], slab["world"],
                                             N.current_stream(phi.device)))
        grid = L.lower_grid(var.nx, var.bcs, slab)  # (slab: coefficient classes follow the GLOBAL plane index)
        # (edge=True on rz meshes: the one-sided face formulas of Laplacian / Grad carry no 1/r term, fdc.py:223-288 --
        #  the rz coefficient tables act on the cells next to the faces only; fixtures tests/golden/rz_edge.pt)
        op, keep = L.lower_op(A_coeffs, nd, phi.dtype, edge=edge_code, dx=var.mesh._dx, adv_const=adv_const,
                              field_device=phi.device)
        code, stream = N.dtype_code(phi.dtype), N.current_stream(phi.device)
        if self.op_type == "Grad":
            out = torch.empty((1, nd, *var.nx), dtype=phi.dtype, device=phi.device)
            N.check(N.lib().pa_grad_apply(grid, op, code, phi.data_ptr(), out.data_ptr(), stream))
        else:
            eq = N.Equation()
            eq.nops = 1
            eq.ops[0] = op
            out = torch.empty_like(phi)
            N.check(N.lib().pa_stencil_apply(grid, eq, code, phi.data_ptr(), out.data_ptr(), stream))
        del keep
        return out

    def reset(self) -> None:
        self.A_coeffs = None
        self.rhs_adj = None

    def set_config(self, config: DiscretizerConfigType) -> None:
        self._config = config

    def __call__(self, *args):
        if len(args) == 1:
            assert isinstance(args[0], Field), "FDC: only `Field` is allowed for var!"
            var = args[0]
            self.A_coeffs = self.build_A_coeffs(var)
            self.rhs_adj = self.adjust_rhs(var)
            return self.apply(self.A_coeffs, var)
        from pyapes_b200.variables.container import Hess, Jac

        assert isinstance(args[0], (Field, Tensor, float, Jac, Hess)), (
            "FDC: for var_j, Field, Tensor, float, Jac and Hess are allowed!"
        )
        assert isinstance(args[1], Field), "FDC: only `Field` is allowed for var_i!"
        self.A_coeffs = self.build_A_coeffs(args[0], args[1], config=self.config)
        self.rhs_adj = self.adjust_rhs(args[0], args[1], config=self.config)
        self.var_addition = args[0]
        return self.apply(self.A_coeffs, args[1])


class Laplacian(Discretizer):
    _op_type = "Laplacian"

    @staticmethod
    def build_A_coeffs(var: Field) -> L.StarCoeffs:
        """fdc.py:375-423: Neumann/Symmetry faces edit the plane next to them
        (2/3, -2/3, 0 | 0, -2/3, 2/3), everything / dx^2."""
        return L.laplacian_star(var.nx, var.mesh._dx, var.bcs, var().dtype, _rz_x(var))

    @staticmethod
    def adjust_rhs(var: Field) -> Tensor:
        """fdc.py:425-458: Neumann faces add (2/3) V n / dx on the plane next to them."""
        out = torch.zeros_like(var())
        if var.bcs is None:
            return out
        dx = var.dx
        for j in range(var.mesh.dim):
            for bc in var.bcs:
                if bc.bc_type != "neumann" or not _owns_face(var, bc):
                    continue
                pl = _plane(var, bc, 1)
                alpha = torch.zeros_like(out[0][pl])
                if var.mesh.coord_sys == "rz":  # fdc.py:440-448
                    dr = dx[j] if j == 0 else 0.0
                    alpha = torch.nan_to_num(1 / 3 * dr / var.mesh.grid[j][pl], nan=0.0, posinf=0.0, neginf=0.0)
                at_bc = _return_bc_val(bc, var, 0, out[0][pl].shape)
                out[0][pl] += (2 / 3 - alpha) * (at_bc * bc.bc_n_vec[j]) / dx[j]
        return out


class Grad(Discretizer):
    """Central gradient; result shape `(var.dim, mesh.dim, *nx)` (fdc.py:461-502)."""

    _op_type = "Grad"

    @staticmethod
    def build_A_coeffs(var: Field) -> L.StarCoeffs:
        return L.grad_star(var.nx, var.mesh._dx, var.bcs, var().dtype)

    @staticmethod
    def adjust_rhs(var: Field) -> Tensor:
        ones = torch.ones_like(var())
        return _grad_like_rhs(var, ones, ones)


def _check_limiter(config: DivConfigType | None) -> str:
    if config is not None and "limiter" in config:
        return config["limiter"].lower()
    warnings.warn("FDM: no limiter is specified. Use `none` (central difference) as a default.")
    return "none"


def _adv_of(var_j, var_i: Field, limiter: str = "none"):
    """fdc.py:775-792 -> float (constant) or a (1,*nx) tensor.

    Jac-driven advection (fdc.py:639-664,730-735,760-763): for a scalar `var_i` the reference takes
    `adv[n2d[i]]` with i running over var.dim == 1 only, i.e. the FIRST component of the Jacobian on every
    mesh axis -- `div(jac, var)` is bit-identical to `div(jac.x[None], var)` [probed on the reference,
    tests/golden/make_golden_jacdiv.py].  A Hess does not get through the reference either: limiter "upwind"
    raises NotImplementedError (fdc.py:650-653) and "none" dies in adjust_rhs with AttributeError; both are
    reproduced."""
    from pyapes_b200.geometry.basis import n2d_coord
    from pyapes_b200.variables.container import Hess, Jac

    if isinstance(var_j, (float, int)) and not isinstance(var_j, bool):
        return float(var_j)
    if isinstance(var_j, Tensor):
        assert var_j.shape == var_i().shape, "FDC Div: adv shape must match var_i shape"
        return var_j
    if isinstance(var_j, Field):
        return var_j()
    if isinstance(var_j, Jac):
        first = n2d_coord(var_i.mesh.coord_sys)[0]
        return var_j[first].unsqueeze(0)
    if isinstance(var_j, Hess):
        if limiter == "upwind":
            raise NotImplementedError(
                "FDC: Upwind limiter is not implemented for Hessians and Jacobians advection term."
            )
        if any(bc.bc_type in ("neumann", "symmetry") for bc in (var_i.bcs or [])):
            raise IndexError(
                "FDC Div: central scheme with Neumann/Symmetry faces is not usable "
                "(the reference raises IndexError in fdc.py:583-584 as well)"
            )
        raise AttributeError("'Hess' object has no attribute 'x'")
    raise TypeError(f"FDC Div: unsupported advection term {type(var_j).__name__}")


class Div(Discretizer):
    """d(u_j phi)/dx_j, limiter "none" (central), "upwind" (the reference's formula,
    fdc.py:746-772: 2 min(u,0) phi[+1] + 2 max(u,0) phi[-1], no 1/dx) or "upwind_fd"
    (new: the first-order upwind difference the reference's own test intends)."""

    _op_type = "Div"

    @staticmethod
    def build_A_coeffs(var_j, var_i: Field, config: DiscretizerConfigType):
        assert "div" in config, "FDC Div: config should contain 'div' key."
        limiter = _check_limiter(config["div"])
        adv = _adv_of(var_j, var_i, limiter)
        if isinstance(adv, float):
            return L.div_star_const(adv, var_i.nx, var_i.mesh._dx, var_i.bcs, var_i().dtype, limiter, _rz_x(var_i))
        if var_i.mesh.coord_sys == "rz":
            raise NotImplementedError("pyapes_b200: field-valued advection on rz meshes (SURVEY.md §8(f) item 3)")
        if adv.shape[0] != 1:
            adv = adv[0:1]  # scalar var_i uses adv[0] on every axis (fdc.py:735,760)
        return L.div_field(adv.contiguous(), var_i.nx, var_i.mesh._dx, var_i.bcs, limiter)

    @staticmethod
    def adjust_rhs(var_j, var_i: Field, config: DiscretizerConfigType) -> Tensor:
        """fdc.py:666-694."""
        if var_i.bcs is None:
            return torch.zeros_like(var_i())
        assert "div" in config, "FDC Div: config should contain 'div' key."
        limiter = _check_limiter(config["div"])
        adv = _adv_of(var_j, var_i, limiter)
        if isinstance(adv, float):
            adv = torch.ones_like(var_i()) * adv
        if limiter == "none":
            return _grad_like_rhs(var_i, 2.0 * adv, 2.0 * adv)
        if limiter == "upwind":
            z = torch.zeros_like(var_i())
            return _grad_like_rhs(var_i, 2.0 * torch.min(adv, z), 2.0 * torch.max(adv, z))
        if limiter == "upwind_fd":
            return torch.zeros_like(var_i())
        if limiter == "quick":
            raise NotImplementedError("FDC Div: quick scheme is not implemented yet.")
        raise RuntimeError(f"FDC Div: {limiter=} is an unknown limiter type.")


class DiffFlux:
    """Diffusive flux `D_ij d(phi)/dx_j` of a scalar field (fdc.py:820-857): a vector Field built
    from the edge=True Jacobian and a Hess container of diffusion coefficients."""

    @staticmethod
    def __call__(diff, var: Field) -> Field:
        from pyapes_b200.geometry.basis import n2d_coord

        jac = jacobian(var)
        flux = Field("DiffFlux", len(jac), var.mesh, None)
        n2d = n2d_coord(var.mesh.coord_sys)
        for i in range(var.mesh.dim):
            acc = torch.zeros_like(var()[0])
            for j in range(var.mesh.dim):
                acc += diff[n2d[i] + n2d[j]] * jac[n2d[j]]
            flux.set_var_tensor(acc, i)
        return flux


class FDC:
    """Collection of the explicit discretisers; like the reference (fdc.py:860-879) the
    operators are class-level singletons and the constructor pushes `config` into them."""

    div: Div = Div()
    laplacian: Laplacian = Laplacian()
    grad: Grad = Grad()
    diffFlux: DiffFlux = DiffFlux()

    def __init__(self, config: DiscretizerConfigType | None = None):
        self.config = config
        if config is not None:
            for key in config:
                scheme = getattr(self, key, None)
                if isinstance(scheme, Discretizer):
                    scheme.set_config(config)

    def update_config(self, scheme: str, target: str, val: str):
        if self.config is not None:
            self.config.setdefault(scheme, {})[target] = val  # type: ignore[misc]
        else:
            self.config = {scheme: {target: val}}  # type: ignore[misc]
        for key in self.config:
            s = getattr(self, key, None)
            if isinstance(s, Discretizer):
                s.set_config(self.config)


def _edge_gradient(mesh, phi0: Tensor) -> Tensor:
    """edge=True central gradient of a BC-less container field (fdc.py:904-907): `(mesh.dim, *nx)`."""
    box = Field("container", 1, mesh, None)
    box.set_var_tensor(phi0)
    fdc = FDC({"grad": {"edge": True}})
    out = fdc.grad(box)[0]
    fdc.grad.reset()
    FDC({"grad": {"edge": False}})
    return out


def jacobian(var: Field):
    """Jacobian of a scalar field as a `Jac` container (fdc.py:896-914)."""
    from pyapes_b200.geometry.basis import n2d_coord
    from pyapes_b200.variables.container import Jac

    assert var().shape[0] == 1, "Scalar: var must be a scalar field."
    n2d = n2d_coord(var.mesh.coord_sys)
    g = _edge_gradient(var.mesh, var[0])
    return Jac(**{n2d[i]: g[i] for i in range(var.mesh.dim)})


def hessian(var: Field):
    """Hessian as a `Hess` container (upper triangle), gradient of every Jacobian component
    (fdc.py:917-944)."""
    from pyapes_b200.geometry.basis import n2d_coord
    from pyapes_b200.variables.container import Hess

    n2d = n2d_coord(var.mesh.coord_sys)
    d = var.mesh.dim
    g = _edge_gradient(var.mesh, var[0])
    data = {}
    for i in range(d):
        gi = _edge_gradient(var.mesh, g[i].contiguous())
        for j in range(i, d):
            data[n2d[i] + n2d[j]] = gi[j]
    return Hess(**data)
