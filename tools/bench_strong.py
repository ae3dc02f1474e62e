"""Config 5 of BASELINE.json: 3-D Poisson N^3 fp64 CG, Dirichlet, slab-decomposed over the ranks of
one node (strong scaling: fixed global grid).  torchrun --nproc-per-node P tools/bench_strong.py [N] [iters] [reps]
Prints one JSON line per repetition (run-to-run stability) and a summary line."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
warnings.filterwarnings("ignore")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.parallel import SlabMesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, [N, N, N], rank, world, dev) if world > 1 else Mesh(Box[0:1, 0:1, 0:1], None, [N, N, N], dev)
g = torch.Generator().manual_seed(1234 + rank)
rhs = torch.rand((1, *mesh.nx), generator=g, dtype=torch.float64).to(dev)
cfg = {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False, "check_every": iters + (iters & 1)}
def run():
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    s = Solver({"fdm": dict(cfg)}); s.set_eq(FDM().laplacian(1.0, var) == rhs)
    rep = s.solve(); assert rep["itr"] == iters, rep
for _ in range(2): run()
vals = []
for rep in range(reps):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 2], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    glups = float(N) ** 3 * iters / (t.item() * 1e-3) / 1e9
    vals.append(glups)
    if rank == 0:
        print(json.dumps({"case": f"config5 CG {N}^3 strong scaling", "rep": rep, "n_gpus": world, "iters": iters,
                          "ms_per_iter": round(t.item() / iters, 4), "GLUP/s": round(glups, 1),
                          "per_gpu_hbm_frac": round(glups / world * 64e9 / 6541.8e9, 3)}), flush=True)
if rank == 0:
    from pyapes_b200 import _native as _N
    mean = sum(vals) / len(vals)
    print(json.dumps({"summary": f"config5 CG {N}^3 strong scaling", "n_gpus": world, "reps": reps,
                      "GLUP/s_mean": round(mean, 1), "GLUP/s_min": round(min(vals), 1), "GLUP/s_max": round(max(vals), 1),
                      "spread_pct": round(100 * (max(vals) - min(vals)) / mean, 2),
                      "per_gpu_hbm_frac_mean": round(mean / world * 64e9 / 6541.8e9, 3),
                      "halo_exchange": "ncclSend/ncclRecv" if os.environ.get("PA_NO_PEER_HALO") or not _N.lib().pa_p2p_enabled()
                                       or _N.lib().pa_p2p_halo_cap() < 8 * N * N else "peer-memory stores inside phase B"}), flush=True)
if world > 1: dist.destroy_process_group()
