"""Precision and device selectors (reference: pyapes/backend.py:13-94).

`DType("single"|"s"|32)` / `DType("double"|"d"|64)` carries the matching torch dtypes and, like
the reference (backend.py:31,38), switches torch's GLOBAL default dtype — user code and the
reference's own tests rely on freshly created tensors matching the mesh precision.
"""
from __future__ import annotations

import torch

TORCH_DEVICE = ["cpu", "cuda", "mps"]
DTYPE_SINGLE = ["single", "s", 32]
DTYPE_DOUBLE = ["double", "d", 64]

_TABLE = {
    True: (torch.float64, torch.complex128, torch.int64),
    False: (torch.float32, torch.complex64, torch.int32),
}


class DType:
    def __init__(self, precision: str | int = "double"):
        if precision in DTYPE_DOUBLE:
            wide = True
        elif precision in DTYPE_SINGLE:
            wide = False
        else:
            raise ValueError("Invalid precision type!")
        self.precision = precision
        self._float, self._complex, self._int = _TABLE[wide]
        torch.set_default_dtype(self._float)

    float = property(lambda self: self._float)
    int = property(lambda self: self._int)
    complex = property(lambda self: self._complex)
    bool = property(lambda self: torch.bool)

    def __eq__(self, other):
        return isinstance(other, DType) and other._float == self._float

    def __repr__(self):
        return f"(torch.dtype){self.precision}"


class TorchDevice:
    """`cuda` is the product path.  `cpu` objects can be built (host logic, tests) but every
    operator/solver call on them raises: there is no CPU compute path in this package."""

    def __init__(self, device_type: str = "cpu"):
        assert device_type.split(":")[0] in TORCH_DEVICE
        self.device_type = device_type
        self._device = torch.device(device_type.lower())

    @property
    def device(self) -> torch.device:
        return self._device

    def __repr__(self) -> str:
        return f"Device on {self.device}"
