"""Multi-GPU parity: the slab-decomposed CG (P ranks) against the single-GPU CG on the same
global problem.  Run under torchrun:  torchrun --nproc-per-node 2 tools/dist_check.py"""
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
cases = [
    ("dirichlet", [70, 48, 64], ["dirichlet"] * 6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0], 1e-8, 3000),
    ("x-neumann+dirichlet", [64, 40, 64], ["neumann", "dirichlet", "dirichlet", "dirichlet", "dirichlet", "dirichlet"],
     [0.3, 0.0, 0.0, 0.5, 0.0, 0.0], 1e-30, 60),
    ("yz-mixed", [48, 36, 32], ["dirichlet", "dirichlet", "neumann", "symmetry", "dirichlet", "neumann"],
     [0.0, 0.2, 0.5, None, 0.0, -0.1], 1e-30, 40),
]
for variant in (0, 1, 2):
    for name, n, kinds, vals, tol, max_it in cases:
        g = torch.Generator().manual_seed(4321)
        rhs_global = torch.rand(1, *n, generator=g, dtype=torch.float64) - 0.5
        cfg = {"method": "cg", "tol": tol, "max_it": max_it, "report": False, "variant": variant, "use_graph": False}
        mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, n, rank, world, dev)
        var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
        a, b = mesh.slab["goff0"], mesh.slab["goff0"] + mesh.slab["n0_local"]
        rhs_local = rhs_global[:, a:b].contiguous().to(dev)
        s = Solver({"fdm": dict(cfg)})
        s.set_eq(FDM().laplacian(1.0, var) == rhs_local)
        rep = s.solve()
        full = gather_owned(var)
        if rank == 0:
            m1 = Mesh(Box[0:1, 0:1, 0:1], None, n, dev)
            v1 = Field("p", 1, m1, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
            s1 = Solver({"fdm": dict(cfg)})
            s1.set_eq(FDM().laplacian(1.0, v1) == rhs_global.to(dev))
            rep1 = s1.solve()
            ref = v1().cpu()
            err = (full - ref).abs().max().item() / (ref.abs().max().item() + 1e-300)
            good = rep["itr"] == rep1["itr"] and abs(rep["tol"] - rep1["tol"]) <= 1e-10 * max(1.0, rep1["tol"]) and err <= 1e-9
            ok &= good
            print(f"[variant {variant}] {name:22s} P={world} itr={rep['itr']} tol={rep['tol']:.6e} | P=1 itr={rep1['itr']} tol={rep1['tol']:.6e} | rel err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
