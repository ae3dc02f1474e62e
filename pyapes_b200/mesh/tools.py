"""Index regions (reference: pyapes/mesh/tools.py)."""
from __future__ import annotations

from pyapes_b200.geometry.basis import DIR_TO_NUM, SIDE_TO_NUM


def boundary_slicer(dim: int, bcs: list) -> list[slice]:
    """Region every solver vector is written in: `[1:-1]` per axis, open on each side whose
    BC is periodic (tools.py:7-20)."""
    ends: list[list[int | None]] = [[1, -1] for _ in range(dim)]
    for bc in bcs:
        if bc.bc_type == "periodic":
            ends[DIR_TO_NUM[bc.bc_face[0]]][SIDE_TO_NUM[bc.bc_face[1]]] = None
    return [slice(*e) for e in ends]


def inner_slicer(dim: int, pad: int | None = 1) -> list[slice]:
    return [slice(pad, -pad if isinstance(pad, int) else None) for _ in range(dim)]
