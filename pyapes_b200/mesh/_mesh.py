"""Equidistant rectangular mesh (reference: pyapes/mesh/_mesh.py:19-318).

Same constructor and attributes as the reference.  Differences, all about memory at the
sizes this package targets (512^3 and up):
  * face masks (`d_mask`, `t_mask`) are built on first access instead of in `__init__`
    (6 full-size bool tensors per mesh in 3-D otherwise);
  * node coordinates are computed once on the host with `torch.linspace` and copied, so the
    Neumann `x_face - x_inner` distances are the bits the reference's CPU path produces.
The CUDA kernels never read the masks or the grid: boundary geometry is passed as indices.
"""
from __future__ import annotations

from functools import cached_property
from typing import Optional

import torch
from torch import Tensor

from pyapes_b200.backend import DTYPE_DOUBLE, DTYPE_SINGLE, TORCH_DEVICE, DType, TorchDevice
from pyapes_b200.geometry import GeoTypeIdentifier
from pyapes_b200.geometry.basis import DIR_TO_NUM, DIR_TO_NUM_RZ, Geometry


class _LazyMasks(dict):
    """dict face -> bool tensor, filled on demand; iteration order = the reference's
    face order (geometry/basis.py:152-199)."""

    def __init__(self, mesh: "Mesh"):
        super().__init__()
        self._mesh = mesh
        self._faces = [str(c["face"]) for c in mesh.domain.config.values()]

    def __missing__(self, face: str) -> Tensor:
        if face not in self._faces:
            raise KeyError(face)
        m = self._mesh
        axis = (DIR_TO_NUM_RZ if m.coord_sys == "rz" else DIR_TO_NUM)[face[0]]
        mask = torch.zeros(*m.nx, dtype=torch.bool, device=m.device)
        mask.select(axis, 0 if face[1] == "l" else m.nx[axis] - 1).fill_(True)
        self[face] = mask
        return mask

    def _fill(self):
        for f in self._faces:
            self[f]

    def __iter__(self):
        self._fill()
        return iter(self._faces)

    def __len__(self):
        return len(self._faces)

    def __contains__(self, face):
        return face in self._faces

    def keys(self):
        return list(self._faces)

    def values(self):
        return [self[f] for f in self._faces]

    def items(self):
        return [(f, self[f]) for f in self._faces]


class Mesh:
    def __init__(
        self,
        domain: Geometry,
        obstacle: Optional[list[Geometry]],
        spacing: list[int] | list[float] = [],
        device: str | None = None,
        dtype: str | int = "double",
    ):
        # The reference's default is "cpu" (mesh/_mesh.py:29).  This package computes on CUDA only, so a
        # script written for the reference's default can be pointed at the GPU without editing it:
        # PYAPES_B200_DEFAULT_DEVICE=cuda replaces the DEFAULT (an explicit device= always wins).
        if device is None:
            import os

            device = os.environ.get("PYAPES_B200_DEFAULT_DEVICE", "cpu")
        # "cuda:N" is accepted as well (one process per GPU in the slab-decomposed runs)
        assert device.split(":")[0] in TORCH_DEVICE, "Mesh: device only accept cpu or cuda"
        self.device = TorchDevice(device).device
        assert dtype in DTYPE_DOUBLE or dtype in DTYPE_SINGLE, "Mesh: dtype only accept double or single"
        self.dtype = DType(dtype)
        self.domain = domain
        if self.coord_sys == "rz":
            assert self.dim == 2, "Mesh: rz coordinate system only accept 2D domain"
        if obstacle is not None:
            raise NotImplementedError(
                "pyapes_b200: inner obstacles are not supported (the reference's solvers raise "
                "NotImplementedError for them as well, linalg.py:287-292)."
            )
        self.obstacle = obstacle

        ft = self.dtype.float
        self._lower = torch.tensor(self.domain.lower, dtype=ft, device=self.device)
        self._upper = torch.tensor(self.domain.upper, dtype=ft, device=self.device)
        self._lx = self._upper - self._lower
        lx_host = [float(v) for v in (torch.tensor(self.domain.upper, dtype=ft) - torch.tensor(self.domain.lower, dtype=ft))]

        if int in GeoTypeIdentifier(spacing):  # number of nodes given
            self._nx: list[int] = [int(s) for s in spacing]
            self._dx: list[float] = [float(torch.tensor(l, dtype=ft) / (n - 1.0)) for l, n in zip(lx_host, self._nx)]
        elif float in GeoTypeIdentifier(spacing):  # spacing given
            self._dx = [float(s) for s in spacing]
            self._nx = [int(torch.tensor(l, dtype=ft) / d + 1.0) for l, d in zip(lx_host, self._dx)]
        else:
            raise TypeError("Mesh: spacing only accept int or float")

        lo_h = torch.tensor(self.domain.lower, dtype=ft)
        up_h = torch.tensor(self.domain.upper, dtype=ft)
        self._x_host = [
            torch.linspace(lo_h[i].item(), up_h[i].item(), self._nx[i], dtype=ft) for i in range(self.dim)
        ]
        self.slab = None
        self._localize()  # hook: a slab-decomposed mesh narrows axis 0 to its own planes here
        self.x = [x.to(self.device) for x in self._x_host]
        # meshgrid returns stride-0 views: no memory is spent here
        self.grid = torch.meshgrid(self.x, indexing="ij")
        self.o_mask: dict = {}

    def _localize(self) -> None:
        """Single-GPU mesh: the local block is the whole grid."""

    # ---- lazily materialised masks ---------------------------------------------------
    @cached_property
    def d_mask(self) -> dict[str, Tensor]:
        return _LazyMasks(self)

    @cached_property
    def t_mask(self) -> Tensor:
        tm = torch.zeros(*self.nx, dtype=torch.bool, device=self.device)
        for axis in range(self.dim):
            tm.select(axis, 0).fill_(True)
            tm.select(axis, self.nx[axis] - 1).fill_(True)
        return tm

    def __repr__(self) -> str:
        return f"{self.domain} with dx={self.dx.tolist()}"

    @property
    def coord_sys(self) -> str:
        if self.domain.type == "box":
            return "xyz"
        if self.domain.type == "cylinder":
            return "rz"
        raise TypeError(f"Mesh: domain type ({self.domain.type=}) not identifiable")

    def d_mask_dim(self, d_face: str) -> int:
        return (DIR_TO_NUM_RZ if self.coord_sys == "rz" else DIR_TO_NUM)[d_face[0]]

    def d_mask_dir(self, d_face: str) -> int:
        return 1 if d_face[1] == "r" else -1

    def d_mask_shift(self, d_face: str, shift: int) -> Tensor:
        return torch.roll(self.d_mask[d_face], -shift * self.d_mask_dir(d_face), self.d_mask_dim(d_face))

    @property
    def _depth(self) -> float:
        if self.dim == 1:
            return self._dx[0] * self._dx[0]
        if self.dim == 2:
            return self._dx[0]
        return 1.0

    dim = property(lambda self: self.domain.dim)
    size = property(lambda self: self.domain.size)
    lx = property(lambda self: self._lx)
    lower = property(lambda self: self._lower)
    upper = property(lambda self: self._upper)
    center = property(lambda self: self._lx * 0.5)
    is_cuda = property(lambda self: self.device.type == "cuda")

    @property
    def N(self) -> int:
        n = 1
        for v in self._nx:
            n *= v
        return n

    @property
    def dx(self) -> Tensor:
        return torch.tensor(self._dx, dtype=self.dtype.float, device=self.device)

    @property
    def nx(self) -> torch.Size:
        return torch.Size(self._nx)

    @property
    def R(self) -> Tensor:
        if self.coord_sys != "rz":
            raise KeyError("Mesh: R coordinate only available in axisymmetric case.")
        return self.grid[0]

    @property
    def X(self) -> Tensor:
        return self.grid[0]

    def _empty(self) -> Tensor:
        return torch.tensor([], dtype=self.dtype.float, device=self.device)

    @property
    def Y(self) -> Tensor:
        if self.coord_sys == "rz":
            return self._empty()
        return self.grid[1] if self.dim > 1 else self._empty()

    @property
    def Z(self) -> Tensor:
        if self.coord_sys == "rz":
            return self.grid[1]
        return self.grid[2] if self.dim > 2 else self._empty()

    @cached_property
    def dg(self) -> list[Tensor]:
        """Half the sum of the forward and backward node distances (one-sided at the ends)."""
        out = []
        for idx, g in enumerate(self.grid):
            fwd = torch.roll(g, -1, idx) - g
            bwd = g - torch.roll(g, 1, idx)
            out.append((fwd.clamp_min(0.0) + bwd.clamp_min(0.0)) / 2)
        return out


def get_box_mask(x: list[Tensor], dx: Tensor, obj: dict, mask: Tensor, dim: int) -> Tensor:
    """Mark the nodes a face description covers (reference: pyapes/mesh/_mesh.py:375-399).

    `obj` is one entry of `Geometry.config`: `x_p` the corner the face starts at, `e_x` its extent
    per axis.  Along every axis the marked range starts at the node nearest to `x_p` and holds
    `ceil(e_x / dx) + 1` nodes."""
    region = []
    for axis in range(dim):
        nodes = x[axis]
        corner = torch.as_tensor(obj["x_p"][axis], dtype=nodes.dtype, device=nodes.device)
        extent = torch.as_tensor(obj["e_x"][axis], dtype=nodes.dtype, device=nodes.device)
        first = int(torch.argmin((nodes - corner).abs()))
        count = int(torch.ceil(extent / dx[axis])) + 1
        region.append(slice(first, first + count))
    mask[tuple(region)] = True
    return mask


def boundary_mask(mesh: Mesh) -> tuple[dict, dict]:
    """`(domain face masks, obstacle masks)` as the reference builds them in `Mesh.__init__`
    (pyapes/mesh/_mesh.py:321-372).  `Mesh.d_mask` holds the same domain masks lazily; obstacles are
    rejected by `Mesh`, so the second dictionary is always empty."""
    faces: dict[str, Tensor] = {}
    for desc in mesh.domain.config.values():
        blank = torch.zeros(*mesh.nx, dtype=torch.bool, device=mesh.device)
        faces[str(desc["face"])] = get_box_mask(mesh.x, mesh.dx, desc, blank, mesh.dim)
    return faces, {}
