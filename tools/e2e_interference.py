"""Does a bulk host<->device copy in flight slow the CG solve down?  (bench.py's e2e trace at N = 8: solves 25 % slower
while the copies of the neighbouring steps run.)  Single process: the 512^3 solve alone, then with pinned-memory copies
looping on two side streams for the whole solve.  Under torchrun: the same on slabs.
   python tools/e2e_interference.py        |   torchrun --nproc-per-node 2 tools/e2e_interference.py"""
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device(dev))
from pyapes_b200.solver.fdm import FDM  # noqa: E402
from pyapes_b200.solver.ops import Solver  # noqa: E402

import bench as B  # noqa: E402

n, iters = 512, 200
cfg = {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False, "check_every": iters}
if world > 1:
    from pyapes_b200.parallel import make_slab_problem

    mesh, var = make_slab_problem(n, rank, world, dev)
else:
    mesh, var = B.make_problem(n, "cuda")
shape = tuple(var().shape)
g = torch.Generator().manual_seed(1234 + rank)
rhs = torch.rand(shape, generator=g, dtype=torch.float64).to(dev)
host_a = torch.empty(1 << 27, dtype=torch.float64).pin_memory()  # 1 GiB
host_b = torch.empty(1 << 27, dtype=torch.float64).pin_memory()
dev_a = torch.empty(1 << 27, dtype=torch.float64, device=dev)
dev_b = torch.empty(1 << 27, dtype=torch.float64, device=dev)
up, down, main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def solve():
    if world > 1:
        m, v = make_slab_problem(n, rank, world, dev)
    else:
        m, v = B.make_problem(n, "cuda")
    s = Solver({"fdm": dict(cfg)})
    s.set_eq(FDM().laplacian(1.0, v) == rhs)
    s.solve()


def timed(copies: str, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if "h2d" in copies:
            with torch.cuda.stream(up):
                for _ in range(12):
                    dev_a.copy_(host_a, non_blocking=True)
        if "d2h" in copies:
            with torch.cuda.stream(down):
                for _ in range(12):
                    host_b.copy_(dev_b, non_blocking=True)
        with torch.cuda.stream(main):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solve()
            e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)[len(out) // 2]


with torch.cuda.stream(main):
    solve()
torch.cuda.synchronize()
res = {"world": world, "rank": rank}
for c in ("none", "h2d", "d2h", "h2d+d2h", "none"):
    res["solve_ms_copies_" + c + ("_again" if c == "none" and "solve_ms_copies_none" in res else "")] = round(timed(c), 2)
if rank == 0:
    print(json.dumps(res), flush=True)
