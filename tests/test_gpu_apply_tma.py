"""The explicit operators on the TMA star engine (PW_APPLY / PW_GRAD, csrc/kernels_tma_pw.cuh) --
`Solver.Aop`, `FDC().laplacian / .grad / .div` (fdc.py:67-118,171-200,461-502,612-694).

* multi-tile / partial-tile fixtures from the REAL reference (ops_tiles.pt): bit-exact, through the
  TMA kernels and through the generic kernels (PA_APPLY_VARIANT=generic);
* larger shapes that reach the predicate-free LEAN tiles: bit-exact against the oracle;
* BASELINE sizes (256^3, 1024^2, 512^3): the TMA kernels against the generic ones, bit-exact
  (the oracle cannot run there in seconds; the generic kernels are pinned by the fixtures).
"""
import os

import pytest
import torch

from tests import _util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
TILES = U.load("ops_tiles.pt")


class _variant:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.prev = os.environ.get("PA_APPLY_VARIANT")
        if self.name in ("generic", "tma"):  # forced paths
            os.environ["PA_APPLY_VARIANT"] = self.name
        else:  # "auto": the direct 2-D kernel on 2-D grids up to 8 M cells, the TMA engine elsewhere
            os.environ.pop("PA_APPLY_VARIANT", None)

    def __exit__(self, *a):
        if self.prev is None:
            os.environ.pop("PA_APPLY_VARIANT", None)
        else:
            os.environ["PA_APPLY_VARIANT"] = self.prev


@pytest.mark.parametrize("variant", ["tma", "generic", "auto"])
@pytest.mark.parametrize("case", TILES, ids=[c["name"] for c in TILES])
def test_tile_fixtures_bit_exact(case, variant):
    mesh, var = U.product_field(case, DEV)
    var.set_var_tensor(case["phi"].to(DEV).clone())
    with _variant(variant):
        got = U.product_tile_outputs(case, var)
    assert sorted(got) == sorted(case["out"])
    for key, ref in case["out"].items():
        g = got[key].cpu()
        assert g.shape == ref.shape, (key, g.shape, ref.shape)
        assert torch.equal(g, ref), f"{key}: max|d|={(g - ref).abs().max().item():.3e}"


def _synthetic_case(nx, kinds, vals, dtype, seed=5):
    nd = len(nx)
    faces = ["xl", "xu", "yl", "yu", "zl", "zu"][: 2 * nd]
    lower, upper = [0.0] * nd, [1.0, 0.75, 1.25][:nd]
    spec = {"lower": lower, "upper": upper, "nx": nx, "dtype": dtype, "bcs": list(zip(kinds, vals))}
    tdt = U.TDTYPE[dtype]
    import oracle.fd_oracle as O

    _, dx = O.make_axes(lower, upper, nx, tdt)
    g = torch.Generator().manual_seed(seed)
    phi = (torch.rand((1, *nx), generator=g, dtype=torch.float64) - 0.5).to(tdt)
    return {"name": "synthetic", "spec": spec, "bcs": list(zip(faces, kinds, vals)), "dx": dx, "phi": phi,
            "u_const": 0.65}


LEAN_SHAPES = [
    ([7, 48, 192], ["dirichlet"] * 6, [0.1, 0, 0, 0.3, 0, 0], "double"),
    ([9, 52, 260], ["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"],
     [None, None, 0.5, None, 0.0, 0.0], "double"),
    ([6, 49, 196], ["neumann", "dirichlet", "dirichlet", "dirichlet", "periodic", "periodic"],
     [0.2, 0.0, 0.0, 1.0, None, None], "double"),
    ([5, 64, 384], ["dirichlet"] * 6, [0.0] * 6, "single"),
    ([37, 2048], ["neumann", "dirichlet", "dirichlet", "dirichlet"], [0.0, 0.0, 1.0, 1.0], "double"),
    ([21, 1540], ["dirichlet", "dirichlet", "periodic", "periodic"], [0.0, 0.5, None, None], "double"),
    ([19, 3080], ["dirichlet"] * 4, [0.0] * 4, "single"),
]


@pytest.mark.parametrize("nx,kinds,vals,dtype", LEAN_SHAPES, ids=[f"{'x'.join(map(str, s[0]))}_{s[3]}" for s in LEAN_SHAPES])
def test_apply_vs_oracle_interior_tiles(nx, kinds, vals, dtype):
    """Shapes with tiles that touch no array edge (the LEAN instantiation) and several chunks."""
    torch.set_default_dtype(U.TDTYPE[dtype])
    case = _synthetic_case(nx, kinds, vals, dtype)
    ref = U.oracle_tile_outputs(case)
    mesh, var = U.product_field(case, DEV)
    var.set_var_tensor(case["phi"].to(DEV).clone())
    got = U.product_tile_outputs(case, var)
    assert sorted(got) == sorted(ref)
    for key, r in ref.items():
        g = got[key].cpu()
        assert torch.equal(g, r), f"{key}: max|d|={(g - r).abs().max().item():.3e}"


FULL = [
    ([256, 256, 256], "double"), ([512, 512, 512], "double"), ([1024, 1024], "double"), ([4096, 4096], "double"),
    ([200, 136, 250], "double"), ([256, 256, 256], "single"), ([130, 1026], "double"),
]


@pytest.mark.parametrize("nx,dtype", FULL, ids=[f"{'x'.join(map(str, s[0]))}_{s[1]}" for s in FULL])
def test_tma_equals_generic_at_baseline_sizes(nx, dtype):
    nd = len(nx)
    kinds = (["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"] if nd == 3
             else ["neumann", "dirichlet", "dirichlet", "dirichlet"])
    vals = [None, None, 0.5, None, 0.0, 0.0] if nd == 3 else [0.0, 0.0, 1.0, 1.0]
    faces = ["xl", "xu", "yl", "yu", "zl", "zu"][: 2 * nd]
    case = {"spec": {"lower": [0.0] * nd, "upper": [1.0] * nd, "nx": nx, "dtype": dtype},
            "bcs": list(zip(faces, kinds, vals)), "u_const": -0.4}
    mesh, var = U.product_field(case, DEV)
    g = torch.Generator().manual_seed(99)
    phi = (torch.rand((1, *nx), generator=g, dtype=torch.float32) - 0.5).to(U.TDTYPE[dtype]).to(DEV)
    var.set_var_tensor(phi)
    with _variant("tma"):
        a = U.product_tile_outputs(case, var)
    with _variant("generic"):
        b = U.product_tile_outputs(case, var)
    for key in b:
        assert torch.equal(a[key], b[key]), f"{key}: max|d|={(a[key] - b[key]).abs().max().item():.3e}"
        del a[key]
    if nd == 2:  # the direct 2-D kernel (auto picks it up to 8 M cells)
        with _variant("auto"):
            c = U.product_tile_outputs(case, var)
        for key in b:
            assert torch.equal(c[key], b[key]), f"direct {key}: max|d|={(c[key] - b[key]).abs().max().item():.3e}"
    assert torch.isfinite(b["lap"]).all()
