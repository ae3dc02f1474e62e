"""Multi-GPU parity: the slab-decomposed solvers (P ranks) against the single-GPU solvers on the
same global problem.  Run under torchrun:  torchrun --nproc-per-node 2 tools/dist_check.py

CG / Jacobi: iteration count exact, tol to 1e-10, solution to 1e-9 relative.
Explicit Euler: bit-exact (no reduction on the path).
BiCGSTAB: fixed-iteration (lockstep) runs agree to 1e-8 relative; converged runs must converge on
both sides (tol 1e-6) to the same solution within 1e-6 relative (the algorithm amplifies the rank-order
difference of the reductions, as the reference does with its own thread count)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
warnings.filterwarnings("ignore")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.parallel import SlabMesh, gather_owned
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import mixed_bcs

ok = True
D6 = ["dirichlet"] * 6
cases = [
    ("dirichlet", [70, 48, 64], D6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0]),
    ("x-neumann+dirichlet", [64, 40, 64], ["neumann"] + ["dirichlet"] * 5, [0.3, 0.0, 0.0, 0.5, 0.0, 0.0]),
    ("yz-mixed", [48, 36, 32], ["dirichlet", "dirichlet", "neumann", "symmetry", "dirichlet", "neumann"],
     [0.0, 0.2, 0.5, None, 0.0, -0.1]),
    ("odd-nz (generic path)", [40, 20, 31], D6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0]),
    # config 4 of BASELINE.json: periodic along the slab axis (ring exchange + cross-rank periodic BC)
    ("x-periodic mixed", [44, 36, 32], ["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"],
     [None, None, 0.5, None, 0.0, 0.0]),
    ("x-periodic, y-periodic", [38, 20, 32], ["periodic"] * 4 + ["dirichlet"] * 2, [None] * 4 + [0.0, 0.3]),
]


def both(build, n, kinds, vals, dtype="double"):
    """Run `build(mesh, var, rhs_local_or_global) -> report` on the slab and (rank 0) on one GPU."""
    tdt = torch.float64 if dtype == "double" else torch.float32
    g = torch.Generator().manual_seed(4321)
    rhs_global = (torch.rand(1, *n, generator=g, dtype=torch.float64) - 0.5).to(tdt)
    x0_global = torch.rand(1, *n, generator=g, dtype=torch.float64).to(tdt)
    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, n, rank, world, dev, dtype, periodic=kinds[0] == "periodic")
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    rep = build(var, mesh.local_slice(rhs_global).to(dev), mesh.local_slice(x0_global).to(dev))
    full = gather_owned(var)
    if rank != 0:
        return None
    m1 = Mesh(Box[0:1, 0:1, 0:1], None, n, dev, dtype)
    v1 = Field("p", 1, m1, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    rep1 = build(v1, rhs_global.to(dev), x0_global.to(dev))
    ref = v1().cpu()
    err = (full - ref).abs().max().item() / (ref.abs().max().item() + 1e-300)
    return rep, rep1, err


def report(tag, name, res, good):
    global ok
    rep, rep1, err = res
    ok &= bool(good)
    print(f"[{tag}] {name:24s} P={world} itr={rep['itr']} tol={rep['tol']:.6e} | P=1 itr={rep1['itr']} "
          f"tol={rep1['tol']:.6e} | rel err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)


def krylov(method, tol, max_it, variant, graph=True):
    def build(var, rhs, x0):
        cfg = {"method": method, "tol": tol, "max_it": max_it, "report": False, "variant": variant, "use_graph": graph}
        s = Solver({"fdm": cfg})
        s.set_eq(FDM().laplacian(1.0, var) == rhs)
        return s.solve()
    return build


# ---- CG: all three kernel variants
for variant in (0, 1, 2):
    for (name, n, kinds, vals), (tol, max_it) in zip(cases, [(1e-8, 3000), (1e-30, 60), (1e-30, 40), (1e-30, 30), (1e-30, 30), (1e-30, 30)]):
        res = both(krylov("cg", tol, max_it, variant, graph=False), n, kinds, vals)
        if res:
            rep, rep1, err = res
            report(f"cg v{variant}", name, res, rep["itr"] == rep1["itr"]
                   and abs(rep["tol"] - rep1["tol"]) <= 1e-10 * max(1.0, rep1["tol"]) and err <= 1e-9)

# ---- CG as the solver runs it by default: iteration pairs replayed as a CUDA graph (fused TMA kernels with the
#      peer-memory halo exchange and in-kernel all-reduces when the mailboxes are up)
from pyapes_b200 import _native as _N
if rank == 0:
    print(f"[info] p2p mailboxes: {bool(_N.lib().pa_p2p_enabled())}, landing plane cap: {_N.lib().pa_p2p_halo_cap()} B "
          f"(0 = ncclSend/ncclRecv halo exchange)", flush=True)
for (name, n, kinds, vals), (tol, max_it) in zip(cases, [(1e-8, 3000), (1e-30, 60), (1e-30, 40), (1e-30, 30), (1e-30, 30), (1e-30, 30)]):
    res = both(krylov("cg", tol, max_it, 0, graph=True), n, kinds, vals)
    if res:
        rep, rep1, err = res
        report("cg v0 graph", name, res, rep["itr"] == rep1["itr"]
               and abs(rep["tol"] - rep1["tol"]) <= 1e-10 * max(1.0, rep1["tol"]) and err <= 1e-9)
res = both(krylov("cg", 1e-9, 4000, 0, graph=True), [96, 80, 192], D6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0])
if res:
    rep, rep1, err = res
    report("cg v0 graph", "dirichlet 96x80x192 (15 tiles per plane)", res, rep["itr"] == rep1["itr"] and rep["converge"]
           and abs(rep["tol"] - rep1["tol"]) <= 1e-10 * max(1.0, rep1["tol"]) and err <= 1e-9)

# ---- Jacobi: TMA star engine (variant 0) and generic kernels (variant 1)
for variant in (0, 1):
    for name, n, kinds, vals in cases:
        res = both(krylov("jacobi", 1e-30, 50, variant), n, kinds, vals)
        if res:
            rep, rep1, err = res
            report(f"jacobi v{variant}", name, res, rep["itr"] == rep1["itr"] == 51
                   and abs(rep["tol"] - rep1["tol"]) <= 1e-10 * max(1.0, rep1["tol"]) and err <= 1e-9)

# ---- BiCGSTAB: lockstep (fixed 12 iterations) and converged
for variant in (0, 1):
    for name, n, kinds, vals in cases:
        res = both(krylov("bicgstab", 1e-30, 12, variant), n, kinds, vals)
        if res:
            rep, rep1, err = res
            report(f"bicgstab v{variant} lockstep", name, res, rep["itr"] == rep1["itr"] == 12
                   and abs(rep["tol"] - rep1["tol"]) <= 1e-7 * max(1.0, rep1["tol"]) and err <= 1e-8)
    # periodic faces: the reference's own converged solutions spread by ~1e-3
    # (the all-Dirichlet case is left out: the reference recursion breaks down on it for some
    # summation orders -- the CPU oracle stalls at 1e-3 with 1e-8 requested -- on one GPU as well)
    for (name, n, kinds, vals), etol in zip(cases[1:3] + [cases[4]], [1e-6, 1e-6, 1e-2]):
        # 1e-6: inside the accuracy the reference's recursion attains on these grids (at 1e-8 the
        # reference algorithm itself stagnates on the first case, on one GPU and in the CPU oracle)
        res = both(krylov("bicgstab", 1e-6, 3000, variant), n, kinds, vals)
        if res:
            rep, rep1, err = res
            report(f"bicgstab v{variant} converged", name, res, rep["converge"] and rep1["converge"]
                   and rep["tol"] <= 1e-6 and abs(rep["itr"] - rep1["itr"]) <= 0.25 * rep1["itr"] and err <= etol)
# fp32 on the slab: lockstep BiCGSTAB and CG
for method in ("bicgstab", "cg"):
    res = both(krylov(method, 1e-30, 12, 0), *cases[0][1:], dtype="single")
    if res:
        rep, rep1, err = res
        report(f"{method} fp32 lockstep", cases[0][0], res, rep["itr"] == rep1["itr"] and err <= 1e-3)


# ---- explicit Euler: upwind Div + Laplacian, bit-exact
def euler(limiter, steps):
    def build(var, rhs, x0):
        var.set_var_tensor(x0.clone())
        fdm = FDM({"div": {"limiter": limiter, "edge": False}})
        nu = 0.1
        dt = 0.2 * min(var.mesh._dx) ** 2 / nu
        var.set_time(dt, 0.0)
        s = Solver({"fdm": {"method": "euler", "n_steps": steps, "report": False}})
        s.set_eq(fdm.ddt(var) + fdm.div(1.0, var) - fdm.laplacian(nu, var) == 0.0)
        return s.solve()
    return build


for limiter in ("upwind", "upwind_fd"):
    for name, n, kinds, vals in cases:
        res = both(euler(limiter, 25), n, kinds, vals)
        if res:
            rep, rep1, err = res
            report(f"euler {limiter}", name, res, err == 0.0)


# ---- nonlinear advection div(var, var) on slabs (fdm.py:306-312): the iterate's ghost planes follow every update
def burgers(method, limiter, max_it):
    def build(var, rhs, x0):
        var.set_var_tensor(0.2 * x0.clone())
        fdm = FDM({"div": {"limiter": limiter, "edge": False}})
        cfg = {"method": method, "tol": 1e-30, "max_it": max_it, "report": False}
        if method == "euler":
            var.set_time(0.05 * min(var.mesh._dx) ** 2 / 0.1, 0.0)
            s = Solver({"fdm": {"method": "euler", "n_steps": max_it, "report": False}})
            s.set_eq(fdm.ddt(var) + fdm.div(var, var) - fdm.laplacian(0.1, var) == 0.0)
        else:
            s = Solver({"fdm": cfg})
            s.set_eq(fdm.div(var, var) - fdm.laplacian(0.1, var) == rhs)
        return s.solve()
    return build


for method, limiter, its, tol in (("bicgstab", "none", 8, 1e-8), ("bicgstab", "upwind", 8, 1e-8), ("jacobi", "upwind", 20, 1e-12),
                                  ("cg", "none", 6, 1e-8), ("euler", "none", 15, 0.0), ("euler", "upwind", 15, 0.0)):
    name, n, kinds, vals = cases[0]
    res = both(burgers(method, limiter, its), n, kinds, vals)
    if res:
        rep, rep1, err = res
        report(f"nonlinear {method}/{limiter}", name, res, rep["itr"] == rep1["itr"] and err <= tol)


# ---- explicit FDC operators on slabs: every owned cell equals the single-GPU value (edge=True: everywhere;
#      edge=False: except the two global x-boundary planes, whose values are wrap-around artefacts)
def fdc_check(kinds, vals, n, periodic):
    global ok
    from pyapes_b200.solver.fdc import FDC
    g = torch.Generator().manual_seed(77)
    phi_global = torch.rand(1, *n, generator=g, dtype=torch.float64) - 0.5
    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, n, rank, world, dev, "double", periodic=periodic)
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    var.set_var_tensor(mesh.local_slice(phi_global).to(dev))
    m1 = Mesh(Box[0:1, 0:1, 0:1], None, n, dev, "double")
    v1 = Field("p", 1, m1, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    v1.set_var_tensor(phi_global.to(dev))
    for opname, edge in (("laplacian", False), ("laplacian", True), ("grad", False), ("grad", True), ("div", False)):
        cfg = {opname: {"edge": edge}} if opname != "div" else {"div": {"limiter": "upwind", "edge": False}}
        call = (lambda f, v: f.div(0.7, v)) if opname == "div" else (lambda f, v: getattr(f, opname)(v))
        loc = call(FDC(cfg), var)
        ref = call(FDC(cfg), v1) if rank == 0 else None
        own = loc[:, :, mesh.slab["olo0"]:mesh.slab["ohi0"]] if opname == "grad" else mesh.owned(loc)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(own.contiguous().cpu(), parts, dst=0)
        if rank == 0:
            full = torch.cat(parts, dim=2 if opname == "grad" else 1)
            r = ref.cpu()
            sl = slice(None) if (edge or periodic) else slice(1, -1)
            a = full[:, :, sl] if opname == "grad" else full[:, sl]
            b = r[:, :, sl] if opname == "grad" else r[:, sl]
            good = torch.equal(a, b)
            ok &= bool(good)
            print(f"[fdc {opname} edge={edge}] {'periodic-x' if periodic else 'dirichlet'} P={world} bit-equal on owned cells: {'OK' if good else 'FAIL'}", flush=True)
    FDC({"laplacian": {"edge": False}, "grad": {"edge": False}, "div": {"limiter": "none", "edge": False}})


fdc_check(D6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0], [40, 36, 64], False)
fdc_check(["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"], [None, None, 0.5, None, 0.0, 0.0], [44, 36, 32], True)

dist.barrier()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
