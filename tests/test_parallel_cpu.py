"""Host-side logic of the slab decomposition, world_size 2 on CPU with the gloo backend
(no CUDA kernels run here: layout, lowering to pa_grid, gather of owned planes)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyapes_b200 import _lower as L
        from pyapes_b200.geometry import Box
        from pyapes_b200.parallel import SlabMesh, gather_owned
        from pyapes_b200.variables import Field
        from pyapes_b200.variables.bcs import mixed_bcs

        n = [11, 5, 6]
        mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, n, rank, world, "cpu")
        kinds = ["neumann", "dirichlet", "dirichlet", "dirichlet", "periodic", "periodic"]
        var = Field("p", 1, mesh, {"domain": mixed_bcs([0.5, 0.0, 0.0, 0.0, None, None], kinds), "obstacle": None})
        s = mesh.slab
        # every owned plane carries its global index
        gidx = torch.arange(s["goff0"], s["goff0"] + s["n0_local"], dtype=torch.float64)
        var.set_var_tensor(gidx.view(1, -1, 1, 1).expand(1, -1, n[1], n[2]).contiguous())
        g = L.lower_grid(mesh.nx, var.bcs, s)
        info = dict(rank=rank, n=list(g.n), lo=list(g.lo), hi=list(g.hi), gn0=g.gn0, goff0=g.goff0, olo0=g.olo0,
                    ohi0=g.ohi0, x0=mesh.x[0].tolist())
        full = gather_owned(var)
        if rank == 0:
            info["gathered_planes"] = full[0, :, 0, 0].tolist()
        out.put(info)
    finally:
        dist.destroy_process_group()


def test_slab_layout_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    infos = sorted((out.get(timeout=120) for _ in procs), key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = infos
    # 11 planes -> 6 + 5; one ghost plane on the interior side of each rank
    assert a["n"] == [7, 5, 6] and b["n"] == [6, 5, 6]
    assert (a["goff0"], a["olo0"], a["ohi0"]) == (0, 0, 6)
    assert (b["goff0"], b["olo0"], b["ohi0"]) == (5, 1, 6)
    assert a["gn0"] == b["gn0"] == 11
    # solver region: owned planes inside the global [1, 10); z is periodic -> open
    assert (a["lo"][0], a["hi"][0]) == (1, 6) and (b["lo"][0], b["hi"][0]) == (1, 5)
    assert a["lo"][2] == 0 and a["hi"][2] == 6 and a["lo"][1] == 1 and a["hi"][1] == 4
    # ghost coordinates are the neighbour's coordinates
    assert a["x0"][-1] == pytest.approx(b["x0"][1]) and b["x0"][0] == pytest.approx(a["x0"][-2])
    assert a["gathered_planes"] == [float(i) for i in range(11)]


def test_partition_covers_everything():
    sys.path.insert(0, ROOT)
    from pyapes_b200.parallel import partition, slab_layout

    for n0, w in ((512, 8), (1024, 8), (100, 7), (9, 3)):
        parts = partition(n0, w)
        assert parts[0][0] == 0 and parts[-1][1] == n0
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in parts) - min(e - s for s, e in parts) <= 1
        lay = [slab_layout(n0, r, w) for r in range(w)]
        assert sum(l["ohi0"] - l["olo0"] for l in lay) == n0
    with pytest.raises(ValueError):
        slab_layout(8, 0, 4)


def test_periodic_slab_layout_and_region():
    """Periodic x faces: every rank carries both ghost planes (ring), rank 0 starts at global -1,
    the solver region covers every owned plane, and local_slice wraps around."""
    sys.path.insert(0, ROOT)
    from pyapes_b200 import _lower as L
    from pyapes_b200.geometry import Box
    from pyapes_b200.parallel import SlabMesh, slab_layout
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    lay = [slab_layout(11, r, 2, periodic=True) for r in range(2)]
    assert [(l["goff0"], l["olo0"], l["ohi0"], l["n0_local"]) for l in lay] == [(-1, 1, 7, 8), (5, 1, 6, 7)]
    assert slab_layout(11, 0, 1, periodic=True)["n0_local"] == 11  # one rank: the kernels wrap themselves
    kinds = ["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"]
    vals = [None, None, 0.5, None, 0.0, 0.0]
    glob = torch.arange(11, dtype=torch.float64).view(1, 11, 1, 1).expand(1, 11, 5, 6)
    for r in range(2):
        mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, [11, 5, 6], r, 2, "cpu", periodic=True)
        var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
        g = L.lower_grid(mesh.nx, var.bcs, mesh.slab)
        assert (g.lo[0], g.hi[0]) == (g.olo0, g.ohi0) == (1, mesh.nx[0] - 1)
        planes = mesh.local_slice(glob)[0, :, 0, 0].tolist()
        assert planes == ([10.0, 0, 1, 2, 3, 4, 5, 6] if r == 0 else [5.0, 6, 7, 8, 9, 10, 0])
        assert mesh.x[0][0].item() == pytest.approx((planes[0]) / 10.0)
    # a periodic Field on a non-periodic slab (or the reverse) is refused when lowering
    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, [11, 5, 6], 0, 2, "cpu")
    var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
    with pytest.raises(ValueError):
        L.lower_grid(mesh.nx, var.bcs, mesh.slab)
