"""Slab decomposition across GPUs (one process per GPU) — new; the reference is single-device
(SURVEY.md §5, §8e).

    torch.distributed.init_process_group("nccl")
    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, [1024, 1024, 1024], rank, world, "cuda")
    var  = Field("p", 1, mesh, {"domain": bcs, "obstacle": None})      # LOCAL slab (+ ghost planes)
    solver.set_eq(fdm.laplacian(1.0, var) == local_rhs); solver.solve()

The global grid is split along mesh axis 0 (the slowest axis of the contiguous `(1,nx,ny,nz)`
layout, so a halo plane is one contiguous run).  Rank p owns planes [start, stop) and stores one
ghost plane per interior side.  `torch.distributed` is only plumbing here (bootstrap of the NCCL
id, test gathers); the halo exchange and the all-reduces of the Krylov scalars run inside the
native solver on a NCCL communicator created from that id (csrc/dist.cuh).
"""
from __future__ import annotations

import ctypes as C

import torch

from pyapes_b200 import _native as N
from pyapes_b200.mesh import Mesh


def partition(n0: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, stop) plane ranges, sizes differing by at most one."""
    base, rem = divmod(n0, world)
    out, s = [], 0
    for p in range(world):
        e = s + base + (1 if p < rem else 0)
        out.append((s, e))
        s = e
    return out


def slab_layout(n0: int, rank: int, world: int, periodic: bool = False) -> dict:
    """Local block of rank `rank`: owned planes [start, stop) plus one ghost plane towards each
    existing neighbour.  `periodic`: axis 0 carries Periodic BCs, so rank 0 and rank world-1 are
    neighbours too and every rank has both ghost planes (the wrap-around pair of SURVEY.md §8e).
    All indices as the C ABI's pa_grid wants them (goff0 is -1 on rank 0 of a periodic slab)."""
    start, stop = partition(n0, world)[rank]
    ring = bool(periodic) and world > 1
    lo_ghost = 1 if (rank > 0 or ring) else 0
    hi_ghost = 1 if (rank < world - 1 or ring) else 0
    if stop - start < 3:
        raise ValueError(f"slab decomposition needs >= 3 planes per rank (n0={n0}, world={world})")
    return {
        "rank": rank, "world": world, "start": start, "stop": stop, "periodic": ring,
        "gn0": n0, "goff0": start - lo_ghost, "olo0": lo_ghost, "ohi0": lo_ghost + (stop - start),
        "n0_local": (stop - start) + lo_ghost + hi_ghost,
    }


class SlabMesh(Mesh):
    """A Mesh whose tensors cover only this rank's slab (owned planes + ghost planes) of the
    global grid.  `nx` is the LOCAL shape; `global_nx` the full one; `slab` the layout.
    Pass `periodic=True` when the Field on this mesh has Periodic BCs on the two x faces."""

    def __init__(self, domain, obstacle, spacing, rank: int, world: int, device: str = "cuda",
                 dtype: str | int = "double", periodic: bool = False):
        self._slab_rank, self._slab_world, self._slab_periodic = int(rank), int(world), bool(periodic)
        super().__init__(domain, obstacle, spacing, device, dtype)

    def _localize(self) -> None:
        if self.dim != 3:
            raise NotImplementedError("SlabMesh: slab decomposition is implemented for 3-D meshes")
        self.global_nx = list(self._nx)
        n0 = self._nx[0]
        self.slab = slab_layout(n0, self._slab_rank, self._slab_world, self._slab_periodic)
        a, b = self.slab["goff0"], self.slab["goff0"] + self.slab["n0_local"]
        idx = torch.tensor([i % n0 for i in range(a, b)], dtype=torch.long)  # wrap-around ghosts
        self._x_host[0] = self._x_host[0][idx].clone()
        self._nx = [self.slab["n0_local"], self._nx[1], self._nx[2]]

    def local_slice(self, t: torch.Tensor) -> torch.Tensor:
        """This rank's block (owned + ghost planes, wrap-around included) of a GLOBAL
        `(1, N0, ny, nz)` tensor."""
        n0 = self.slab["gn0"]
        a, b = self.slab["goff0"], self.slab["goff0"] + self.slab["n0_local"]
        idx = torch.tensor([i % n0 for i in range(a, b)], dtype=torch.long, device=t.device)
        return t.index_select(1, idx).contiguous()

    def owned(self, t: torch.Tensor) -> torch.Tensor:
        """View of the owned planes of a local `(1, n0_local, ny, nz)` tensor."""
        return t[:, self.slab["olo0"]: self.slab["ohi0"]]


_COMM: dict = {}


def get_comm(device) -> C.c_void_p:
    """NCCL communicator of the native library for this process (created on first use).
    Rank 0 makes the NCCL unique id; torch.distributed broadcasts it."""
    import torch.distributed as dist

    if "comm" in _COMM:
        return _COMM["comm"]
    if not dist.is_initialized():
        raise RuntimeError("pyapes_b200.parallel: call torch.distributed.init_process_group first")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = N.lib()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        N.check(lib.pa_comm_unique_id(buf))
    on_gpu = dist.get_backend() == "nccl"
    t = torch.tensor(list(buf), dtype=torch.uint8, device=device if on_gpu else "cpu")
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
    comm = C.c_void_p()
    torch.cuda.set_device(device)
    N.check(lib.pa_comm_create(idbuf, rank, world, C.byref(comm)))
    _COMM["comm"] = comm
    _attach_peer_mailboxes(lib, rank, world, device, on_gpu)
    return comm


def _attach_peer_mailboxes(lib, rank: int, world: int, device, on_gpu: bool) -> None:
    """Exchange the CUDA IPC handles of the per-rank mailboxes so that the CG kernels can all-reduce
    their dot products over NVLink peer memory inside the kernel (csrc/common.cuh p2p_allreduce).
    Falls back to ncclAllReduce (with a warning) if any rank cannot export or map a mailbox."""
    import os
    import warnings

    import torch.distributed as dist

    if world < 2 or os.environ.get("PA_NO_P2P"):
        return
    mine = (C.c_ubyte * 64)()
    ok = lib.pa_p2p_local_handle(mine) == 0
    t = torch.tensor(list(mine) + [1 if ok else 0], dtype=torch.uint8, device=device if on_gpu else "cpu")
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    rows = [bytes(a.cpu().tolist()) for a in allt]
    good = all(r[64] == 1 for r in rows)
    if good:
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(r[:64] for r in rows))
        good = lib.pa_p2p_attach(blob, rank, world) == 0
    # [attached?, bytes per landing plane]: every rank attached or nobody uses the mailboxes; the halo landing
    # zones (csrc/common.cuh HaloDev) are used with the smallest capacity any rank could allocate
    flag = torch.tensor([1 if good else 0, int(lib.pa_p2p_halo_cap())], dtype=torch.int64,
                        device=device if on_gpu else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag[0].item()) != 1:
        lib.pa_p2p_disable()
        warnings.warn("pyapes_b200: peer mailboxes unavailable, CG scalars go through ncclAllReduce")
    else:
        lib.pa_p2p_set_halo_cap(int(flag[1].item()))
    _COMM["p2p"] = int(flag[0].item()) == 1


def destroy_comm() -> None:
    if "comm" in _COMM:
        N.lib().pa_comm_destroy(_COMM.pop("comm"))


def make_slab_problem(n: int, rank: int, world: int, device: str, dtype: str = "double"):
    """Weak-scaling Poisson problem of bench.py: global grid (n*world) x n x n on the unit cube
    scaled along x, homogeneous Dirichlet, every rank holds n^3 owned points."""
    from pyapes_b200.geometry import Box
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    mesh = SlabMesh(Box([0.0, 0.0, 0.0], [float(world), 1.0, 1.0]), None, [n * world, n, n], rank, world, device, dtype)
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    return mesh, var


def gather_owned(var, dst: int = 0):
    """Global `(1, N0, ny, nz)` CPU tensor on rank `dst` (None elsewhere) — test helper."""
    import torch.distributed as dist

    mesh = var.mesh
    local = mesh.owned(var()).contiguous().cpu()
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    return torch.cat(parts, dim=1)
