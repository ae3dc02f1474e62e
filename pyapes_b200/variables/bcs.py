"""Boundary conditions (reference: pyapes/variables/bcs.py).

Same classes, constructor keywords, properties and config helpers as the reference:
`Dirichlet`, `Neumann`, `Symmetry`, `Periodic`, `BC_FACTORY`, `BoxBoundary`,
`CylinderBoundary`, `mixed_bcs`, `homogeneous_bcs`, `BC_HD`, `BC_HN`.

What changed underneath: `apply` launches the CUDA face kernel (csrc/kernels_generic.cuh,
k_bc_face) instead of bool-mask `index_put`; the five full-size masks per BC
(bcs.py:84-93) are materialised only if somebody asks for them (a callable `bc_val` does).
"""
from __future__ import annotations

from typing import Callable, NamedTuple, TypedDict, get_args

import torch
from torch import Tensor

from pyapes_b200 import _lower as L
from pyapes_b200 import _native as N
from pyapes_b200.backend import DType
from pyapes_b200.geometry.basis import DIR_TO_NUM, DIR_TO_NUM_RZ, FDIR, FDIR_RZ

BC_val_type = int | float | list[int] | list[float] | Callable | Tensor | None


class BCConfig(TypedDict):
    bc_face: str
    bc_type: str
    bc_val: BC_val_type
    bc_val_opt: dict[str, Tensor] | None


def _bc_val_type_check(bc_val: BC_val_type):
    if not callable(bc_val) and type(bc_val) not in get_args(BC_val_type):
        raise TypeError(f"BC: wrong bc variable -> {type(bc_val)} is not one of {get_args(BC_val_type)}!")


class BC:
    """One boundary face of one variable.  `bc_mask` may be the bool mask itself (as in the
    reference) or a zero-argument callable producing it on demand."""

    def __init__(self, bc_id: str, bc_val: BC_val_type, bc_val_opt, bc_face: str, bc_mask,
                 bc_var_name: str, bc_coord_sys: str, mesh_dim: int, dtype: DType, device: torch.device):
        _bc_val_type_check(bc_val)
        self.bc_id, self.bc_val, self.bc_val_opt, self.bc_face = bc_id, bc_val, bc_val_opt, bc_face
        self.bc_var_name, self.bc_coord_sys, self.mesh_dim = bc_var_name, bc_coord_sys, mesh_dim
        self.dtype, self.device = dtype, device
        self._mask_src = bc_mask
        table = DIR_TO_NUM_RZ if bc_coord_sys == "rz" else DIR_TO_NUM
        self._bc_face_dim = table[bc_face[0]]
        self._bc_n_dir = -1 if bc_face[-1] == "l" else 1
        self._bc_type = self.__class__.__name__.lower()
        self._rolled: dict[int, Tensor] = {}

    # ---- masks (lazy) ------------------------------------------------------------------
    @property
    def bc_mask(self) -> Tensor:
        if callable(self._mask_src):
            self._mask_src = self._mask_src()
        return self._mask_src

    def bc_mask_shift(self, shift: int) -> Tensor:
        return torch.roll(self.bc_mask, shift, self.bc_face_dim)

    def _roll(self, k: int) -> Tensor:
        if k not in self._rolled:
            self._rolled[k] = torch.roll(self.bc_mask, k, self.bc_face_dim)
        return self._rolled[k]

    bc_mask_prev = property(lambda self: self._roll(-self.bc_n_dir))
    bc_mask_prev2 = property(lambda self: self._roll(-2 * self.bc_n_dir))
    bc_mask_forward = property(lambda self: self._roll(self.bc_n_dir))
    bc_mask_forward2 = property(lambda self: self._roll(2 * self.bc_n_dir))

    @property
    def bc_n_vec(self) -> Tensor:
        v = torch.zeros(3, dtype=self.dtype.float, device=self.device)
        v[self.bc_face_dim] = self.bc_n_dir
        return v

    @property
    def bc_treat(self) -> bool:
        return self.bc_type in ("neumann", "symmetry")

    bc_type = property(lambda self: self._bc_type)
    type = property(lambda self: self._bc_type)
    bc_face_dim = property(lambda self: self._bc_face_dim)
    bc_n_dir = property(lambda self: self._bc_n_dir)

    def __repr__(self):
        return f"{self.__class__.__name__}(bc_id={self.bc_id!r}, bc_face={self.bc_face!r}, bc_val={self.bc_val!r})"

    # ---- the seam: BC.apply(var, grid, var_dim) (bcs.py:185-194) ---------------------------
    def apply(self, var: Tensor, grid: tuple[Tensor, ...], var_dim: int) -> None:
        """In place on `var[var_dim]`, through the CUDA face kernel."""
        assert grid
        N.require_cuda(var, "variable")
        nd = var.dim() - 1
        comp = var[var_dim]
        face, keep = L.lower_face(self, grid, var, var_dim, nd)
        g = L.lower_grid(comp.shape, [])
        arr = (N.FaceBC * 1)(face)
        N.check(N.lib().pa_bc_apply(g, 1, arr, N.dtype_code(var.dtype), comp.data_ptr(),
                                    N.current_stream(var.device)))
        del keep


class Dirichlet(BC):
    """phi[face] = value (bcs.py:197-213)."""


class Neumann(BC):
    """Second-order one-sided: phi[face] = 4/3 phi[1] - 1/3 phi[2] + 2/3 V (x_f - x_1) n
    (bcs.py:216-253)."""


class Symmetry(BC):
    """phi[face] = phi[1] (bcs.py:256-262)."""


class Periodic(BC):
    """lower: phi[0] = phi[1] - phi[N-1] + phi[N-2]; upper: phi[N-1] = phi[0] (bcs.py:265-280)."""


class BCContainer(TypedDict, total=False):
    bc_type: str
    bc_val: BC_val_type
    bc_val_opt: dict[str, Tensor] | None


def _get_bc_dict(bc_config, fdir: list[str]) -> list[BCConfig]:
    out: list[BCConfig] = []
    for face in fdir:
        d = getattr(bc_config, face)
        if d is not None:
            out.append({"bc_face": face, "bc_type": d["bc_type"], "bc_val": d["bc_val"],
                        "bc_val_opt": d["bc_val_opt"] if "bc_val_opt" in d else None})
    return out


class CylinderBoundary(NamedTuple):
    rl: BCContainer | None = None
    ru: BCContainer | None = None
    zl: BCContainer | None = None
    zu: BCContainer | None = None

    def __call__(self) -> list[BCConfig]:
        return _get_bc_dict(self, FDIR_RZ)


class BoxBoundary(NamedTuple):
    xl: BCContainer | None = None
    xu: BCContainer | None = None
    yl: BCContainer | None = None
    yu: BCContainer | None = None
    zl: BCContainer | None = None
    zu: BCContainer | None = None

    def __call__(self) -> list[BCConfig]:
        return _get_bc_dict(self, FDIR)


def mixed_bcs(bc_val: list[BC_val_type], bc_type: list[str]) -> list[BCConfig]:
    """Faces in FDIR order: xl, xu, yl, yu, zl, zu (bcs.py:385-408)."""
    return [{"bc_face": FDIR[i], "bc_type": t, "bc_val": v, "bc_val_opt": None}
            for i, (v, t) in enumerate(zip(bc_val, bc_type))]


def homogeneous_bcs(dim: int, bc_val, bc_type: str) -> list[BCConfig]:
    return [{"bc_face": FDIR[i], "bc_type": bc_type,
             "bc_val": bc_val[i] if isinstance(bc_val, list) else bc_val, "bc_val_opt": None}
            for i in range(dim * 2)]


class BC_HD:
    def __new__(cls, dim: int, bc_val: float):
        return homogeneous_bcs(dim, bc_val, "dirichlet")


class BC_HN:
    def __new__(cls, dim: int, bc_val: float):
        return homogeneous_bcs(dim, bc_val, "neumann")


BC_type = Dirichlet | Neumann | Symmetry | Periodic

BC_FACTORY: dict[str, type[BC]] = {
    "dirichlet": Dirichlet,
    "neumann": Neumann,
    "symmetry": Symmetry,
    "periodic": Periodic,
}
