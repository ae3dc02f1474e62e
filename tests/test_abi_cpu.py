"""CPU checks of the drop-in boundary: the library loads without a GPU, exports every symbol
that include/pyapes_b200.h declares, the ctypes structs match the header's layout, and compute
entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pyapes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G

    G.build()
    from pyapes_b200 import _native as N

    lib = N.lib()
    names = _header_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pyapes_b200.h but not exported"
        assert name in N.SYMBOLS, f"{name} has no ctypes prototype in pyapes_b200/_native.py"
    assert lib.pa_abi_version() == 2


def test_struct_sizes_match_header():
    from pyapes_b200 import _native as N

    assert C.sizeof(N.FaceBC) == 32
    assert C.sizeof(N.Grid) == 60
    assert C.sizeof(N.Op) == 4 + 4 + 8 + 8 + 27 * 8 + 8 + 24 + 24 + 12 + 12 + 8 + 4 + 4 + 8 + 24
    assert C.sizeof(N.Equation) == 8 + 4 * C.sizeof(N.Op)
    assert C.sizeof(N.Report) == 32 and C.sizeof(N.SolverCfg) == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from pyapes_b200 import _native as N

    lib = N.lib()
    assert lib.pa_device_count() == 0
    g, eq = N.Grid(), N.Equation()
    rc = lib.pa_stencil_apply(g, eq, N.PA_F64, None, None, None)
    assert rc == -2 and b"no CPU path" in lib.pa_last_error()
    with pytest.raises(N.NativeError):
        N.check(rc)


def test_host_layer_rejects_cpu_fields():
    from pyapes_b200._native import NativeError
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdc import FDC
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    mesh = Mesh(Box[0:1, 0:1], None, [8, 8], "cpu")
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(2, 0.0, "dirichlet"), "obstacle": None})
    s = Solver({"fdm": {"method": "cg", "tol": 1e-6, "max_it": 10, "report": False}})
    s.set_eq(FDM().laplacian(var) == 1.0)
    for call in (s.solve, lambda: s.Aop(var), lambda: FDC({"grad": {"edge": False}}).grad(var),
                 lambda: var.bcs[0].apply(var(), mesh.grid, 0)):
        with pytest.raises(NativeError):
            call()


def test_dsl_semantics_match_reference():
    """Equation algebra (fdm.py:75-105) and set_eq's in-place RHS mutation (ops.py:74-81)."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    mesh = Mesh(Box[0:1, 0:1], None, [9, 7], "cpu")
    bcs = mixed_bcs([0.5, 0.0, 1.0, 0.0], ["neumann", "dirichlet", "neumann", "dirichlet"])
    var = Field("p", 1, mesh, {"domain": bcs, "obstacle": None})
    fdm = FDM({"div": {"limiter": "upwind", "edge": False}})
    rhs = torch.zeros_like(var())
    s = Solver({"fdm": {"method": "bicgstab", "tol": 1e-6, "max_it": 10, "report": False}})
    eq = fdm.div(0.5, var) - fdm.laplacian(0.1, var)
    s.set_eq(eq == rhs)
    assert [s.eqs[k]["name"] for k in s.eqs] == ["Div", "Laplacian"]
    assert s.eqs[0]["sign"] == 1.0 and s.eqs[1]["sign"] == -1
    assert s.rhs is rhs and rhs.abs().sum() > 0  # Neumann adjustment added in place
    assert fdm.div.ops == {} and fdm.div.rhs is None  # singletons reset
    with pytest.raises(RuntimeError):
        s.config["fdm"]["method"] = "gmres"
        from pyapes_b200.solver.linalg import solve

        solve(var, rhs, None, s.eqs, s.config["fdm"], mesh)
    from pyapes_b200.geometry import Cylinder

    rz = Mesh(Cylinder[0:1, 0:2], None, [5, 7], "cpu")
    assert rz.coord_sys == "rz" and list(rz.d_mask) == ["zl", "zu", "rl", "ru"] and rz.R.shape == (5, 7)
    with pytest.raises(KeyError):
        mesh.R


def test_derivative_containers():
    """tests/test_spatial.py:81-128 of the reference."""
    from pyapes_b200.variables.container import Hess, Jac

    x, y, z = torch.rand(10), torch.rand(10), torch.rand(10)
    j = Jac(x=x)
    assert len(j) == 1 and j.keys == ["x"]
    j = Jac(x=x, y=y, z=z)
    assert len(j) == 3 and all(torch.equal(a, b) for a, b in zip(j, [x, y, z]))
    j = Jac(r=x, z=y)
    assert len(j) == 2 and sorted(j.keys) == ["r", "z"]
    h = Hess(xx=x, yy=y)
    assert len(h) == 2 and all(torch.equal(a, b) for a, b in zip(h, [x, y]))
    h = Hess(xx=x, xy=x, xz=x, yy=y, yz=y, zz=z)
    assert all(torch.equal(a, b) for a, b in zip(h, [x, x, x, y, y, z]))
    assert h["zx"] is h.xz
    with pytest.raises(KeyError):
        Hess(rr=x, zz=z)["xx"]


def test_mask_helpers_and_burgers_fixture():
    """Module-level helpers of the reference that user code imports next to `Mesh`
    (pyapes/mesh/_mesh.py:321-399, pyapes/testing/burgers.py, pyapes/geometry/box.py:9)."""
    from math import pi

    from pyapes_b200.geometry import Box
    from pyapes_b200.geometry.box import BOX_DIM
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.mesh._mesh import boundary_mask, get_box_mask
    from pyapes_b200.testing.burgers import burger_exact_nd

    assert BOX_DIM == [1, 2, 3]
    for box, spacing in [(Box[0:1], [11]), (Box[0:1, 0:2], [7, 9]), (Box[0:1, 0:1, 0:1], [0.1, 0.1, 0.25])]:
        mesh = Mesh(box, None, spacing, "cpu", "double")
        faces, obstacles = boundary_mask(mesh)
        assert obstacles == {} and list(faces) == list(mesh.d_mask)
        for name, mask in faces.items():
            assert mask.dtype == torch.bool and torch.equal(mask, mesh.d_mask[name])
            assert int(mask.sum()) == mesh.N // mesh.nx[mesh.d_mask_dim(name)]
    # an interior block: nearest node to the corner, ceil(extent / dx) + 1 nodes per axis
    mesh = Mesh(Box[0:1, 0:1], None, [11, 11], "cpu", "double")
    blank = torch.zeros(11, 11, dtype=torch.bool)
    got = get_box_mask(mesh.x, mesh.dx, {"x_p": [0.2, 0.5], "e_x": [0.25, 0.0], "face": "q"}, blank, 2)
    want = torch.zeros(11, 11, dtype=torch.bool)
    want[2:6, 5:6] = True
    assert torch.equal(got, want)
    # Burgers: the exact profile is 2*pi-periodic in the two-Gaussian approximation and equals 4 where
    # phi_x vanishes (x = pi at t = 0)
    mesh = Mesh(Box[0 : 2 * pi], None, [101], "cpu", "double")
    u0 = burger_exact_nd(mesh, 0.1, 0.0)
    assert u0.shape == mesh.X.shape and abs(u0[50].item() - 4.0) < 1e-12
    assert abs(u0[0].item() - u0[-1].item()) < 1e-12
    with pytest.raises(NotImplementedError):
        burger_exact_nd(Mesh(Box[0:1, 0:1], None, [5, 5], "cpu", "double"), 0.1, 0.0)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/pyapes_b200.h must compile as C99 (and as C++) on its own, and a C
    program must link against the library using nothing but that header."""
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("needs gcc")
    lib_dir = os.path.join(ROOT, "pyapes_b200", "lib")
    src = (
        '#include "pyapes_b200.h"\n'
        "#include <stdio.h>\n"
        "int main(void) {\n"
        "  pa_grid g; pa_report r; r.itr = 0; (void)g;\n"
        '  const char* e = pa_last_error();\n'
        '  printf("%d %d\\n", (int)sizeof(pa_grid), (int)(e != 0) + r.itr);\n'
        "  return 0;\n"
        "}\n"
    )
    with tempfile.TemporaryDirectory() as tmp:
        c_file = os.path.join(tmp, "use_header.c")
        with open(c_file, "w") as f:
            f.write(src)
        inc = os.path.join(ROOT, "include")
        subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", c_file],
                       check=True)
        subprocess.run([shutil.which("g++") or gcc, "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++",
                        c_file], check=True)
        exe = os.path.join(tmp, "use_header")
        subprocess.run([gcc, "-std=c99", "-I", inc, c_file, "-o", exe, "-L", lib_dir, "-lpyapes_b200",
                        f"-Wl,-rpath,{lib_dir}"], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
        from pyapes_b200 import _native as N

        assert int(out[0]) == C.sizeof(N.Grid)


def test_traffic_profile_is_stamped_with_the_current_cg_sources():
    """bench.py refuses `roofline.traffic` (null, "stale") when profiles/ncu_traffic.json was captured on other sources
    of the fused CG kernels than the ones in the tree.  This tripwire turns that into a red CPU test: after touching
    kernels_tma.cuh (or what it includes) re-run tools/prof_cg.py under `ncu --set full` and tools/ncu_traffic.py."""
    import json
    import os

    import __graft_entry__ as G

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        pytest.skip("no profiles/ncu_traffic.json")
    with open(path) as f:
        d = json.load(f)
    assert d["source_hash"] == G._cg_kernel_hash(), "profiles/ncu_traffic.json is stale: re-capture the CG kernels"
    assert 5.3e9 < d["phaseB_512"] < 6.0e9 and 3.2e9 < d["phaseA_512"] < 3.6e9  # 40 / 24 B per cell + halo re-reads
