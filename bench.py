#!/usr/bin/env python3
"""Benchmark of the hot path: 3-D fp64 matrix-free CG on a Poisson problem (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--iters 200] [--impl reference]

A *step* is one `solver.solve()` of `--iters` CG iterations (tol = 1e-30 so the count is
fixed) on an n^3 Dirichlet Poisson problem with a seeded random RHS.
LUP = one grid point x one CG iteration.  Weak scaling: every GPU holds n^3 points
(global grid (n*P) x n x n split in slabs along axis 0).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
public API with the RHS in pinned host memory and the solution copied back, copies inside the
timed region; `roofline` is for the dominant kernel (CG phase B, 5 words/cell) from CUDA events
around every launch of a separate instrumented pass; `cpu_baseline` times the oracle port of the
reference's torch CPU algorithm on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GLUP/s & % HBM roofline, 3D fp64 Laplacian+CG 512^3, 1/2/4/8 B200"
B_PER_LUP_CG = 64.0       # SURVEY.md §8d canonical CG: 8 words x 8 B
B_PER_CELL_PHASE_B = 40.0  # R x, R d, R r, W x, W r
B_PER_CELL_PHASE_A = 24.0  # R r, R d, W d


def workload_name(n, iters):
    return (f"3D Poisson {n}^3 per GPU fp64 matrix-free CG, Dirichlet BCs, {iters} iterations per step "
            f"(tol=1e-30, max_it={iters - 1}); RHS torch.rand seed 1234+rank")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_median": statistics.median(pw) if pw else None}


def make_problem(n, device, rank=0, world=1, dtype="double"):
    """Per-rank problem (single GPU: the whole n^3 box)."""
    import torch

    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    mesh = Mesh(Box[0:1, 0:1, 0:1], None, [n, n, n], device, dtype)
    var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    return mesh, var


def host_rhs(n, rank=0, pin=True):
    import torch

    g = torch.Generator().manual_seed(1234 + rank)
    rhs = torch.rand(1, n, n, n, generator=g, dtype=torch.float64)
    return rhs.pin_memory() if pin and torch.cuda.is_available() else rhs


_CPU_THREADS = None


def cpu_threads():
    """Thread count of the CPU arm: every core this process may run on (torchrun pins
    OMP_NUM_THREADS=1, so it is set explicitly) -- unless fewer threads are FASTER on this host
    (a container whose CPU quota is below its visible core count makes the OpenMP team thrash:
    measured here, 8 visible cores, 30x slower with 8 threads than with 1).  Picked once by timing
    the oracle's stencil application on a 64^3 sample at {all, 1/2, 1/4, 1} of the visible cores."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import torch

    from oracle import fd_oracle as O

    try:
        avail = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        avail = max(1, os.cpu_count() or 1)
    n = 64
    xs, dx = O.make_axes([0, 0, 0], [1, 1, 1], [n] * 3)
    bcs = [O.FaceBC(f, "dirichlet", 0.0) for f in O.FACES]
    phi = torch.rand(1, n, n, n, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(phi)
    best, best_t = avail, float("inf")
    for thr in sorted({avail, max(1, avail // 2), max(1, avail // 4), 1}, reverse=True):
        torch.set_num_threads(thr)
        eq.aop(phi)
        t0 = time.perf_counter()
        for _ in range(2):
            eq.aop(phi)
        dt = time.perf_counter() - t0
        if dt < 0.9 * best_t:  # prefer more threads unless fewer are clearly faster
            best, best_t = thr, dt
    _CPU_THREADS = best
    return best


def cpu_baseline(n_cpu=256, iters=20):
    """Oracle port of the reference's CPU torch algorithm (roll + full coefficient tensors),
    all usable host threads (cpu_threads), bounded sample.  Returns (GLUP/s, seconds, threads, itr)."""
    import torch

    from oracle import fd_oracle as O

    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(cpu_threads())
    xs, dx = O.make_axes([0, 0, 0], [1, 1, 1], [n_cpu] * 3)
    bcs = [O.FaceBC(f, "dirichlet", 0.0) for f in O.FACES]
    x0 = torch.zeros(1, n_cpu, n_cpu, n_cpu, dtype=torch.float64)
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, n_cpu, n_cpu, n_cpu, generator=g, dtype=torch.float64)
    eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(x0)
    eq.adjust_rhs(x0, rhs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        _, rep, _ = O.cg(eq, x0, rhs, 1e-30, iters - 1)
        dt = time.perf_counter() - t0
    lups = n_cpu**3 * rep["itr"] / dt
    return lups / 1e9, dt, torch.get_num_threads(), rep["itr"]


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the reference is pure
    Python so there is no oracle/_ref) on the host cores, same metric/unit/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    n_cpu, iters = args.cpu_n, args.cpu_iters
    for _ in range(max(args.warmup, 0) and 1):
        cpu_baseline(32, 3)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, thr, it = cpu_baseline(n_cpu, iters)
        vals.append(v)
        secs.append(dt)
    value = sum(vals) / len(vals)
    sample = f"{n_cpu}^3 Dirichlet Poisson, {iters} CG iterations per step, torch CPU fp64"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GLUP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.iters), "timed_on": f"bounded CPU sample: {sample}"},
        "cpu_baseline": {"value": value, "unit": "GLUP/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pyapes_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(dev))
    import __graft_entry__ as G

    G.build()
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver

    n, iters = args.n, args.iters
    cfg = {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False, "check_every": iters + (iters & 1)}
    if world > 1:
        from pyapes_b200.parallel import slab_layout

        lay = slab_layout(n * world, rank, world)
        local_shape, olo, ohi = (1, lay["n0_local"], n, n), lay["olo0"], lay["ohi0"]
    else:
        local_shape, olo, ohi = (1, n, n, n), 0, n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def new_solver(rhs_dev):
        if world > 1:
            from pyapes_b200.parallel import make_slab_problem

            mesh, var = make_slab_problem(n, rank, world, dev)
        else:
            mesh, var = make_problem(n, "cuda")
        solver = Solver({"fdm": dict(cfg)})
        solver.set_eq(FDM().laplacian(1.0, var) == rhs_dev)
        return solver, var

    rhs_h = host_rhs(n, rank)  # this rank's OWNED planes (pinned host memory)
    out_h = torch.empty_like(rhs_h).pin_memory()
    rhs_d = torch.zeros(local_shape, dtype=torch.float64, device=dev)
    rhs_d[:, olo:ohi].copy_(rhs_h)
    launches = 0

    def step_device():
        nonlocal launches
        solver, var = new_solver(rhs_d)  # Dirichlet: adjust_rhs adds zeros, rhs_d is unchanged
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rep = solver.solve()
        assert rep["itr"] == iters, rep
        launches += solver.var._last_launches if hasattr(solver.var, "_last_launches") else 0
        return var

    # End-to-end: every step's RHS starts in pinned host memory and its solution ends there.  The
    # copies run on two copy streams so that step k+1's H2D and step k-1's D2H overlap step k's
    # solve (double-buffered device RHS); all of them are inside the timed region.
    # (compute runs on its own stream too: torch's default stream is the legacy stream, which
    # implicitly synchronises with every blocking stream and would serialise the copies.)
    h2d_s, d2h_s, main_s = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    rhs_buf = [torch.zeros(local_shape, dtype=torch.float64, device=dev) for _ in range(2)]

    def run_e2e(k_steps):
        torch.cuda.synchronize()
        with torch.cuda.stream(main_s):
            _run_e2e(k_steps)
        torch.cuda.synchronize()

    def _run_e2e(k_steps):
        ev_in = [None, None]

        def upload(k):
            with torch.cuda.stream(h2d_s):
                rhs_buf[k % 2][:, olo:ohi].copy_(rhs_h, non_blocking=True)
                ev_in[k % 2] = torch.cuda.Event()
                ev_in[k % 2].record(h2d_s)

        upload(0)
        for k in range(k_steps):
            torch.cuda.current_stream().wait_event(ev_in[k % 2])
            if k + 1 < k_steps:
                upload(k + 1)  # overlaps this step's solve
            solver, var = new_solver(rhs_buf[k % 2])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                solver.solve()  # returns when the device has finished this solve
            done = torch.cuda.Event()
            done.record()
            sol = var()
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(done)
                out_h.copy_(sol[:, olo:ohi], non_blocking=True)  # overlaps the next solve
            sol.record_stream(d2h_s)  # the allocator may recycle it only after the download
            del solver, var, sol
        torch.cuda.current_stream().wait_stream(d2h_s)
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    # --- device-resident timing -----------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_timed = launches
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    lup_total = float(n) ** 3 * iters * args.steps * world
    value = lup_total / (ms * 1e-3) / 1e9

    # --- end to end (host buffers) ------------------------------------------------------------
    run_e2e(2)
    # PCIe rate of the two copies alone (explains the fill/drain share of e2e; not part of any metric)
    c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    with torch.cuda.stream(h2d_s):
        c0.record()
        rhs_buf[0][:, olo:ohi].copy_(rhs_h, non_blocking=True)
        c1.record()
        out_h.copy_(rhs_buf[0][:, olo:ohi], non_blocking=True)
        c2.record()
    torch.cuda.synchronize()
    gib = rhs_h.numel() * 8 / 1e9
    copy_rates = {"h2d_GBps": gib / (c0.elapsed_time(c1) * 1e-3), "d2h_GBps": gib / (c1.elapsed_time(c2) * 1e-3)}
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    clocks_e2e = sampler2.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = lup_total / (t.item() * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- kernel-level roofline: every launch bracketed by CUDA events on the launching stream --
    from pyapes_b200 import profile as P

    hbm, peak_src = peaks()
    kt = P.cg_kernel_times(n, iters=20)
    cells = float(n) ** 3
    ach_b = B_PER_CELL_PHASE_B * cells / (kt["phaseB_ms"] * 1e-3) / 1e9
    ach_a = B_PER_CELL_PHASE_A * cells / (kt["phaseA_ms"] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(f"phaseB_{n}")

    cpu_v, cpu_s, cpu_thr, cpu_it = cpu_baseline(args.cpu_n, args.cpu_iters)
    out = {
        "metric": METRIC, "value": value, "unit": "GLUP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(n, iters),
            "global_grid": [n * world, n, n], "decomposition": f"slab x{world} along axis 0",
            "l2_policy": f"working set {8 * 7 * cells / 2**30:.1f} GiB per GPU >> 126 MB L2 (inputs larger than L2)",
            "lup_definition": "grid points x CG iterations",
            "algorithmic_bytes_per_lup": B_PER_LUP_CG,
            "step_hbm_frac": (B_PER_LUP_CG * value / world) / hbm,
        },
        "roofline": {"bound": "hbm", "kernel": kt["kernels"][1], "achieved": ach_b, "peak": hbm, "unit": "GB/s",
                     "frac": ach_b / hbm, "traffic": traffic, "peak_source": peak_src,
                     "avg_launch_ms": kt["phaseB_ms"], "algorithmic_bytes_per_launch": B_PER_CELL_PHASE_B * cells,
                     "other_kernels": {kt["kernels"][0]: {"avg_launch_ms": kt["phaseA_ms"], "achieved": ach_a,
                                                               "frac": ach_a / hbm},
                                       "bc_faces+shell_norm_ms_per_iter": kt["small_ms"]},
                     "kernel_share_of_iteration": kt["share"]},
        "cpu_baseline": {"value": cpu_v, "unit": "GLUP/s", "cores": cpu_thr, "kind": "port",
                         "sample": f"{args.cpu_n}^3 Dirichlet Poisson, {cpu_it} CG iterations, oracle (torch CPU fp64), {cpu_s:.1f} s"},
        "e2e": {"value": e2e_value, "unit": "GLUP/s", "h2d_bytes_per_step": int(rhs_h.numel() * 8),
                "d2h_bytes_per_step": int(out_h.numel() * 8), "clocks": clocks_e2e,
                "pinned_copy_rates": copy_rates,
                "note": "copies of step k+1 / k-1 overlap the solve of step k; the first upload and the last "
                        "download are exposed"},
        "gpu_launches": int(launches_timed),
        "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=512, help="grid points per axis per GPU")
    # the throughput run is a fixed-count solve with max_it = 200
    ap.add_argument("--iters", type=int, default=200, help="CG iterations per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-n", type=int, default=256, help="grid of the bounded CPU sample (reference arm)")
    ap.add_argument("--cpu-iters", type=int, default=20, help="CG iterations of the CPU sample (~10-30 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
