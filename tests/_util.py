"""Shared helpers for the parity tests: load golden fixtures, build oracle objects."""
from __future__ import annotations

import os

import torch

from oracle import fd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TDTYPE = {"double": torch.float64, "single": torch.float32}


def load(name: str):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def oracle_bcs(case) -> list[O.FaceBC]:
    return [O.FaceBC(f, k, v) for f, k, v in case["bcs"]]


def oracle_axes(case):
    spec = case["spec"]
    xs, dx = O.make_axes(spec["lower"], spec["upper"], spec["nx"], TDTYPE[spec["dtype"]])
    assert dx == case["dx"], (dx, case["dx"])
    return xs, dx


def case_rhs(case, shape, dtype):
    rhs = case["rhs"]
    if isinstance(rhs, tuple) and rhs[0] == "rand":
        g = torch.Generator().manual_seed(rhs[1])
        return torch.rand(shape, generator=g, dtype=torch.float64).to(dtype)
    if isinstance(rhs, torch.Tensor):
        return rhs.clone()
    return torch.zeros(shape, dtype=dtype) + float(rhs)


def oracle_terms(case):
    limiter = (case.get("div_cfg") or {}).get("div", {}).get("limiter", "none")
    return [O.Term(kind, sign=float(sign), param=param, limiter=limiter if kind == "div" else "none")
            for kind, sign, param in case["terms"]]
