"""`Field`: a `(dim, *nx)` tensor with boundary conditions and a time stamp
(reference: pyapes/variables/fields.py:19-422).  Same constructor, properties and operator
sugar.  `copy()` / `zeros_like()` share the Mesh instead of deep-copying it (the reference's
deepcopy costs 63 B/point in 3-D)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch
from torch import Tensor

from pyapes_b200.mesh import Mesh
from pyapes_b200.variables.bcs import BC_FACTORY, BC_type, BCConfig


@dataclass
class Field:
    name: str
    dim: int
    mesh: Mesh
    bc_config: dict[str, list[BCConfig] | None] | None
    init_val: Any = None
    object_interp: bool = False

    def __post_init__(self):
        self._VAR = torch.zeros(self.dim, *self.mesh.nx, dtype=self.mesh.dtype.float,
                                device=self.mesh.device, requires_grad=False)
        iv = self.init_val
        if iv is not None:
            if isinstance(iv, float):
                self._VAR += iv
            elif isinstance(iv, list):
                assert self.dim == len(iv), "Field: init_val should match with Field dimension!"
                if not isinstance(iv[0], (float, Tensor)):
                    raise ValueError(f"Field: {type(iv[0])} is an unsupported init_val type!")
                for d in range(self.dim):
                    self._VAR[d] += iv[d] if isinstance(iv[d], Tensor) else float(iv[d])
            elif isinstance(iv, Tensor):
                assert self.dim == iv.size(0), "Field: init_val should match with Field dimension!"
                for d in range(self.dim):
                    self._VAR[d] += iv[d]
            elif isinstance(iv, str) and iv.lower() == "random":
                self._VAR = torch.rand_like(self._VAR)
            else:
                raise ValueError("Field: unsupported data type!")
        if self.bc_config is not None:
            if "domain" not in self.bc_config:
                raise ValueError("Field: domain must be defined!")
            if "obstacle" not in self.bc_config:
                self.bc_config["obstacle"] = None
        self.set_bcs()

    # ---- geometry / time -------------------------------------------------------------
    @property
    def mesh_axis(self) -> list[int]:
        return [i + 1 for i in range(self.mesh.dim)]

    def set_time(self, dt: float, init_val: float | None = None) -> None:
        self._t = init_val if init_val is not None else 0.0
        self._dt = dt

    def update_time(self, dt: float | None = None) -> None:
        self._t += self.dt if dt is None else dt

    t = property(lambda self: self._t)
    dt = property(lambda self: self._dt)
    dx = property(lambda self: self.mesh.dx)
    nx = property(lambda self: self.mesh.nx)

    # ---- storage -----------------------------------------------------------------------
    def save_old(self) -> None:
        self._VARo = self._VAR.clone()

    @property
    def VARo(self) -> Tensor:
        return self._VARo

    @VARo.setter
    def VARo(self, other: Tensor) -> None:
        self._VARo = other

    @property
    def VAR(self) -> Tensor:
        return self._VAR

    @VAR.setter
    def VAR(self, other: Tensor) -> None:
        self._VAR = other

    def _clone_shell(self, name: str | None, tensor: Tensor) -> "Field":
        new = object.__new__(Field)
        new.__dict__.update(self.__dict__)
        new._VAR = tensor
        if "_VARo" in new.__dict__:
            new._VARo = self._VARo.clone()
        new.name = self.name if name is None else name
        new.set_bcs()
        return new

    def copy(self, name: str | None = None) -> "Field":
        return self._clone_shell(name, self._VAR.clone())

    def zeros_like(self, name: str | None = None) -> "Field":
        return self._clone_shell(name, torch.zeros_like(self._VAR))

    def zeros_like_tensor(self) -> Tensor:
        return torch.zeros_like(self._VAR)

    @property
    def size(self) -> torch.Size:
        return self._VAR.size()

    def sum(self, dim: int = 0) -> Tensor:
        return torch.sum(self._VAR, dim=dim)

    def set_var_tensor(self, val: Tensor, insert: int | None = None) -> "Field":
        """Rebinds the storage when shapes match (fields.py:226-227), else broadcasts `val`
        into every component (or only component `insert`)."""
        if self.size == val.shape:
            self._VAR = val
        else:
            for i in range(self.dim):
                if insert is None or i == insert:
                    self._VAR[i] = val
        return self

    def __getitem__(self, idx: int | slice) -> Tensor:
        return self._VAR if isinstance(idx, slice) else self._VAR[idx]

    def __setitem__(self, idx: int | slice, val: Tensor) -> None:
        if isinstance(idx, slice):
            self._VAR = val
        else:
            self._VAR[idx] = val

    def __call__(self) -> Tensor:
        return self._VAR

    # ---- in-place operator sugar (fields.py:256-337) --------------------------------------
    def __add__(self, other: Any) -> "Field":
        if isinstance(other, Field):
            self._VAR += other()
        elif isinstance(other, float):
            self._VAR += other
        elif isinstance(other, list):
            assert len(other) == self.dim, "Field: input vector should match with Field dimension!"
            for i in range(self.dim):
                self._VAR[i] += other[i]
        elif isinstance(other, Tensor):
            if other.size(0) == self.dim:
                self._VAR = other
            else:
                for i in range(other.size(0)):
                    self._VAR[i] += other[i]
        else:
            raise TypeError("Field: you can only add Field, float, Tensor, list[int], or list[float]!")
        return self

    def __sub__(self, other: Any) -> "Field":
        if not isinstance(other, Field):
            raise TypeError("Field: you can only subtract Field!")
        self._VAR -= other()
        return self

    def __mul__(self, other: Any) -> "Field":
        if isinstance(other, Field):
            self._VAR *= other()
        elif isinstance(other, (float, int)):
            self._VAR *= other
        else:
            raise TypeError("Field: you can only multiply Field, int, or float!")
        return self

    def __truediv__(self, other: Any) -> "Field":
        if not isinstance(other, Field):
            raise TypeError("Field: you can only divide by Field!")
        mask = other().gt(0.0)
        self._VAR[mask] /= other()[mask]
        return self

    def __ilshift__(self, other: Any) -> "Field":
        if isinstance(other, Field):
            self._VAR = other()
        elif isinstance(other, Tensor):
            self.set_var_tensor(other)
        elif isinstance(other, (float, int)):
            self._VAR = torch.zeros_like(self._VAR) + other
        elif isinstance(other, list):
            assert self.dim == len(other), "Field: dimension mismatch!"
            self._VAR = torch.zeros_like(self._VAR)
            for i in range(self.dim):
                self._VAR[i] += other[i]
        else:
            raise TypeError("Field: you can only assign Field, Tensor, float, int, or list!")
        return self

    def volume_integral(self, target: Tensor | None = None) -> Tensor:
        if target is None:
            target = torch.ones_like(self._VAR[0])
        val = torch.zeros(self.dim, device=self._VAR.device, dtype=self._VAR.dtype)
        for i in range(self.dim):
            if self.mesh.coord_sys == "xyz":
                val[i] = torch.sum(target * self._VAR[i] * self.mesh.dx.prod())
            else:  # axisymmetric: 2 pi r dr dz (fields.py:350-356)
                val[i] = torch.sum(2.0 * torch.pi * self._VAR[i] * self.mesh.grid[0] * self.mesh.dx.prod())
        return val

    # ---- boundary conditions ---------------------------------------------------------------
    def get_bc(self, bc_id: str) -> BC_type | None:
        found = [bc for bc in self.bcs if bc.bc_id == bc_id]
        if len(found) > 1:
            raise KeyError(f"Field: bc_id {bc_id} returned multiple bcs. Check id once again!")
        return found[0] if found else None

    def set_bcs(self) -> None:
        """BC objects in the order of the config list (that order is the order faces are
        applied in, bcs.py:363-440 / linalg.py:295-297)."""
        self.bcs: list[BC_type] = []
        if self.bc_config is None or self.bc_config["domain"] is None:
            return
        d_bc = self.bc_config["domain"]
        n_faces = len(self.mesh.domain.config)
        assert n_faces == len(d_bc), f"Field: domain config ({n_faces}) mismatch with bc config ({len(d_bc)})!"
        mesh = self.mesh
        for bc in d_bc:
            face = bc["bc_face"]
            self.bcs.append(
                BC_FACTORY[str(bc["bc_type"])](
                    bc_id=f"d-{face}",
                    bc_val=bc["bc_val"],
                    bc_val_opt=bc["bc_val_opt"] if "bc_val_opt" in bc else None,
                    bc_face=face,
                    bc_mask=(lambda f=face: mesh.d_mask[f]),
                    bc_var_name=self.name,
                    bc_coord_sys=mesh.coord_sys,
                    mesh_dim=mesh.dim,
                    dtype=mesh.dtype,
                    device=mesh.device,
                )
            )
        if mesh.obstacle is not None and self.bc_config["obstacle"] is not None:
            raise NotImplementedError
