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
around every launch of a separate instrumented pass; `cpu_baseline` times the reference's CPU
path (the real reference from baseline/_ref or /root/reference when present, else the oracle port)
on a bounded sample.  Outside the headline timed region the same line carries
  `secondary`: the other BASELINE.json configs (2: CG 256^3; 3: Euler 1024^2 / 256^3, both limiters;
               4: BiCGSTAB / Jacobi 512^3 mixed BCs; at N > 1, 5: CG 1024^3 strong scaling), the three
               explicit operator applications and fp32 CG -- flat keys <case>_glups / _hbm_frac /
               _ms / _sm_mhz (words per LUP: SURVEY.md §8d);
  `parity`:    the GPU's 256^3 x 20-iteration CG against the CPU run `cpu_baseline` performs anyway
               (itr, tol <= 1e-10, solution <= 1e-9 relative) and, at N > 1, a P-rank slab solve
               against the single-GPU solve of the same small problem (`dist_ok`).
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


def kernel_source_hash():
    """Content hash of the sources the headline CG kernels are compiled from (kernels_tma.cuh, what it includes, their
    translation units, the nvcc flags)."""
    import __graft_entry__ as G

    return G._cg_kernel_hash()


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu
    capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep together with the hash
    of the kernel sources it was taken on).  A figure whose hash is not the current sources' is REFUSED (null):
    the profile is stale."""
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(tp):
        return None, "no profiles/ncu_traffic.json"
    with open(tp) as f:
        d = json.load(f)
    if d.get("source_hash") != kernel_source_hash():
        return None, f"stale: captured on sources {str(d.get('source_hash'))[:12]}, current {kernel_source_hash()[:12]}"
    return d.get(key), f"ncu --set full capture {d.get('capture', '?')} on sources {d['source_hash'][:12]}"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,timestamp")

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

    @staticmethod
    def _when(cell):
        import datetime

        try:
            return datetime.datetime.strptime(cell, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def window(self, t0, t1):
        """Median SM clock of the samples whose nvidia-smi timestamp lies in [t0, t1] (time.time())."""
        v = []
        for r in list(self.rows):
            try:
                w = self._when(r[7])
                if w is not None and t0 - 0.05 <= w <= t1 + 0.05 and float(r[0]) > 0:
                    v.append(float(r[0]))
            except Exception:
                pass
        return statistics.median(v) if v else None

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
    """Thread count of the CPU arm: every core this process may run on (torchrun pins OMP_NUM_THREADS=1, so
    it is set explicitly).  One exception, decided by a 64^3 probe of the stencil application: if ONE thread
    is at least 1.5x faster than all of them the host's CPU quota is below its visible core count (the
    OpenMP team thrashes: 30x slower with 8 threads than with 1 in the build container) and one thread is
    used.  No other value is ever picked, so the count does not float from run to run."""
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
    t = {}
    for thr in (avail, 1):
        torch.set_num_threads(thr)
        eq.aop(phi)
        t0 = time.perf_counter()
        for _ in range(2):
            eq.aop(phi)
        t[thr] = time.perf_counter() - t0
    _CPU_THREADS = 1 if (avail > 1 and 1.5 * t[1] < t[avail]) else avail
    return _CPU_THREADS


def load_reference():
    """The REAL reference package (pure Python + torch) if it can be found: /root/reference (build
    container) or baseline/_ref (travels to the GPU box, __graft_entry__.install_reference), with the
    3-line stand-in for its un-vendored dependency pymytools.indices (tests/golden/_shim).  None otherwise."""
    for root in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(root, "pyapes")):
            shim = os.path.join(ROOT, "tests", "golden", "_shim")
            for q in (shim, root):
                if q not in sys.path:
                    sys.path.insert(0, q)
            try:
                import pyapes  # noqa: F401

                return root
            except Exception as e:  # noqa: BLE001
                print(f"[bench] reference at {root} failed to import: {e}", file=sys.stderr)
    return None


def cpu_baseline(n_cpu=256, iters=20, keep=False, prefer_reference=True):
    """The reference's CPU torch path on a bounded sample: n_cpu^3 Dirichlet Poisson, seeded RHS, `iters` CG
    iterations, all usable host threads (cpu_threads).  Runs the REAL reference through its own public API
    when load_reference() finds it (kind "reference"), else the oracle port (kind "port").
    Returns dict(value GLUP/s, seconds, threads, itr, tol, kind, solution if keep)."""
    import torch

    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(cpu_threads())
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, n_cpu, n_cpu, n_cpu, generator=g, dtype=torch.float64)
    root = load_reference() if prefer_reference else None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if root is not None:
            from pyapes.geometry import Box
            from pyapes.mesh import Mesh
            from pyapes.solver.fdm import FDM
            from pyapes.solver.ops import Solver
            from pyapes.variables import Field
            from pyapes.variables.bcs import homogeneous_bcs

            mesh = Mesh(Box[0:1, 0:1, 0:1], None, [n_cpu] * 3, "cpu", "double")
            var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
            solver = Solver({"fdm": {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False}})
            solver.set_eq(FDM().laplacian(1.0, var) == rhs)
            t0 = time.perf_counter()
            rep = solver.solve()
            dt = time.perf_counter() - t0
            sol = var()
            kind = "reference"
        else:
            from oracle import fd_oracle as O

            xs, dx = O.make_axes([0, 0, 0], [1, 1, 1], [n_cpu] * 3)
            bcs = [O.FaceBC(f, "dirichlet", 0.0) for f in O.FACES]
            x0 = torch.zeros(1, n_cpu, n_cpu, n_cpu, dtype=torch.float64)
            eq = O.Equation([O.Term("laplacian", 1.0, 1.0)], dx, xs, bcs).build(x0)
            eq.adjust_rhs(x0, rhs)
            t0 = time.perf_counter()
            sol, rep, _ = O.cg(eq, x0, rhs, 1e-30, iters - 1)
            dt = time.perf_counter() - t0
            kind = "port"
    out = {"value": n_cpu**3 * rep["itr"] / dt / 1e9, "seconds": dt, "threads": torch.get_num_threads(),
           "itr": int(rep["itr"]), "tol": float(rep["tol"]), "kind": kind,
           "sample": f"{n_cpu}^3 Dirichlet Poisson, {int(rep['itr'])} CG iterations, "
                     + ("the reference's own solver.solve() (torch CPU fp64)" if kind == "reference"
                        else "oracle port (torch CPU fp64)") + f", {dt:.1f} s"}
    if keep:
        out["solution"] = sol
    return out


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores -- the real
    reference (baseline/_ref or /root/reference) through its public API, else the oracle port -- same
    metric / unit / config, each step a bounded sample.  ONE host process whatever --gpus says: under
    torchrun rank 0 alone runs it, the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu, iters = args.cpu_n, args.cpu_iters
    if args.warmup > 0:
        cpu_baseline(32, 3)
    runs = [cpu_baseline(n_cpu, iters) for _ in range(args.steps)]
    value = sum(r["value"] for r in runs) / len(runs)
    secs = sum(r["seconds"] for r in runs) / len(runs)
    kind, thr = runs[0]["kind"], runs[0]["threads"]
    sample = runs[0]["sample"].rsplit(",", 1)[0]
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GLUP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.iters), "timed_on": f"bounded CPU sample: {sample}",
                   "host_processes": 1,
                   "note": "one host process on rank 0 whatever n_gpus is: the reference has no distributed path"},
        "cpu_baseline": {"value": value, "unit": "GLUP/s", "cores": thr, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def parity_vs_cpu(cpu, n_cpu, iters, dev):
    """(i) of `parity`: the GPU's lockstep CG (same seeded RHS, same fixed count) against the CPU run that
    cpu_baseline has just performed -- config-2 size when --cpu-n is 256."""
    import torch

    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver

    mesh, var = make_problem(n_cpu, dev)
    g = torch.Generator().manual_seed(1234)
    rhs = torch.rand(1, n_cpu, n_cpu, n_cpu, generator=g, dtype=torch.float64).to(dev)
    solver = Solver({"fdm": {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False}})
    solver.set_eq(FDM().laplacian(1.0, var) == rhs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rep = solver.solve()
    ref = cpu["solution"]
    got = var().cpu()
    scale = ref.abs().max().item() or 1.0
    sol_err = (got - ref).abs().max().item() / scale
    tol_diff = abs(rep["tol"] - cpu["tol"])
    ok = rep["itr"] == cpu["itr"] and tol_diff <= 1e-10 and sol_err <= 1e-9
    return {"cpu_kind": cpu["kind"], "case": f"CG {n_cpu}^3 Dirichlet, {iters} iterations (lockstep), RHS seed 1234",
            "itr_gpu": int(rep["itr"]), "itr_cpu": int(cpu["itr"]), "tol_gpu": float(rep["tol"]),
            "tol_cpu": float(cpu["tol"]), "tol_abs_diff": tol_diff, "solution_rel_err": sol_err,
            "bar": "itr equal, |tol diff| <= 1e-10, solution <= 1e-9 relative", "ok": bool(ok)}


def parity_dist(rank, world, dev):
    """(ii) of `parity`, N > 1: a small slab-decomposed CG solve (converged, tol 1e-8) on all ranks against
    the single-GPU solve of the same global problem on rank 0: iteration count, tol, solution."""
    import torch

    from pyapes_b200.geometry import Box
    from pyapes_b200.mesh import Mesh
    from pyapes_b200.parallel import SlabMesh, gather_owned
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import mixed_bcs

    n = [12 * world + 5, 48, 64]
    kinds, vals = ["dirichlet"] * 6, [0.0, 1.0, 0.5, 0.0, -0.25, 0.0]
    g = torch.Generator().manual_seed(4321)
    rhs_global = torch.rand(1, *n, generator=g, dtype=torch.float64) - 0.5
    cfg = {"method": "cg", "tol": 1e-8, "max_it": 3000, "report": False}

    def solve(mesh, rhs):
        var = Field("p", 1, mesh, {"domain": mixed_bcs(vals, kinds), "obstacle": None})
        s = Solver({"fdm": dict(cfg)})
        s.set_eq(FDM().laplacian(1.0, var) == rhs)
        return s.solve(), var

    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, n, rank, world, dev)
    rep, var = solve(mesh, mesh.local_slice(rhs_global).to(dev))
    full = gather_owned(var)
    if rank != 0:
        return None
    rep1, v1 = solve(Mesh(Box[0:1, 0:1, 0:1], None, n, dev), rhs_global.to(dev))
    ref = v1().cpu()
    err = (full - ref).abs().max().item() / (ref.abs().max().item() or 1.0)
    ok = rep["itr"] == rep1["itr"] and abs(rep["tol"] - rep1["tol"]) <= 1e-10 and err <= 1e-9
    return {"dist_case": f"CG {'x'.join(map(str, n))} Dirichlet tol 1e-8, {world} slabs vs 1 GPU",
            "dist_itr": int(rep["itr"]), "dist_itr_1gpu": int(rep1["itr"]),
            "dist_tol_abs_diff": abs(rep["tol"] - rep1["tol"]), "dist_solution_rel_err": err, "dist_ok": bool(ok)}


def secondary_single(hbm, sampler):
    """The other single-GPU configs of BASELINE.json, the explicit operators and fp32 CG, each as a short
    fixed-count run outside the headline's timed region.  Flat dict: <case>_glups, _hbm_frac, _ms, _sm_mhz."""
    import torch

    from pyapes_b200 import profile as P

    out = {}

    def put(tag, r):
        out[f"{tag}_glups"] = round(r["GLUP/s"], 2)
        out[f"{tag}_hbm_frac"] = round(r["GB/s"] / hbm, 4)
        out[f"{tag}_ms"] = round(r["ms"], 4)
        out[f"{tag}_words_per_lup"] = r.get("words_per_lup", r.get("words_per_cell"))
        clk = sampler.window(r["t0"], r["t1"]) if sampler is not None else None
        if clk is not None:
            out[f"{tag}_sm_mhz"] = clk
        if tag.startswith("cfg3_euler"):
            # the adv-diff Euler step is bound by the fp64 ISSUE rate, not by HBM: one rounding per reference operation
            # means 41 (3-D) / 29 (2-D) non-fusable DADD / DMUL per cell; the pipe issues 64 lanes per SM and clock
            ops = 41 if "256" in tag else 29
            out[f"{tag}_fp64_ops_per_lup"] = ops
            out[f"{tag}_fp64_issue_frac"] = round(r["GLUP/s"] * 1e9 * ops / (148 * 64 * 1965e6), 4)

    def timed(fn):
        t0 = time.time()
        r = fn()
        torch.cuda.synchronize()
        r["t0"], r["t1"] = t0, time.time()
        return r

    D4, D6 = (["dirichlet"] * 4, [0.0] * 4), (["dirichlet"] * 6, [0.0] * 6)
    MIX = P.MIXED_BCS
    cases = [
        ("cfg2_cg_256", lambda: P.solver_throughput([256] * 3, "cg", 200, *D6, reps=3, warm=3)),  # (first case after the CPU leg: clocks)
        ("cfg3_euler_1024sq_upwind", lambda: P.euler_throughput([1024, 1024], "upwind", 2000, warm=3)),
        ("cfg3_euler_1024sq_upwind_fd", lambda: P.euler_throughput([1024, 1024], "upwind_fd", 2000, warm=3)),
        ("cfg3_euler_256_upwind", lambda: P.euler_throughput([256] * 3, "upwind", 400)),
        ("cfg3_euler_256_upwind_fd", lambda: P.euler_throughput([256] * 3, "upwind_fd", 400)),
        ("cfg4_bicgstab_512_mixed", lambda: P.solver_throughput([512] * 3, "bicgstab", 100, *MIX)),
        ("cfg4_jacobi_512_mixed", lambda: P.solver_throughput([512] * 3, "jacobi", 100, *MIX)),
        ("cg_1024sq", lambda: P.solver_throughput([1024, 1024], "cg", 1000, *D4, warm=3)),
        ("cg_512_fp32", lambda: P.solver_throughput([512] * 3, "cg", 200, *D6, dtype="single", reps=3, warm=2)),
        # opt-in FMA contraction (PA_FLAG_CONTRACT): the headline solve with ~40 % fewer fp64 instructions
        ("cg_512_contract", lambda: P.solver_throughput([512] * 3, "cg", 200, *D6, reps=3, contract=True)),
        ("cg_512_exact_same_run", lambda: P.solver_throughput([512] * 3, "cg", 200, *D6, reps=3)),
        ("op_laplacian_512", lambda: P.operator_apply_times([512] * 3, "laplacian", reps=40)),
        ("op_grad_512", lambda: P.operator_apply_times([512] * 3, "grad", reps=20)),
        ("op_div_upwind_512", lambda: P.operator_apply_times([512] * 3, "div_upwind", reps=40)),
        ("op_laplacian_256", lambda: P.operator_apply_times([256] * 3, "laplacian", reps=200)),
        ("op_grad_256", lambda: P.operator_apply_times([256] * 3, "grad", reps=100)),
        ("op_laplacian_1024sq", lambda: P.operator_apply_times([1024, 1024], "laplacian", reps=400)),
    ]
    for tag, fn in cases:
        try:
            put(tag, timed(fn))
        except Exception as e:  # noqa: BLE001  (one failing case must not lose the line)
            out[f"{tag}_error"] = f"{type(e).__name__}: {e}"[:200]
        torch.set_default_dtype(torch.float64)
        torch.cuda.empty_cache()
    out["note"] = ("fixed-count runs through the public API, CUDA events; hbm_frac = GLUP/s x words x element size / "
                   "measured HBM peak; words per LUP from SURVEY.md 8d (BiCGSTAB: the canonical 17); 1024^2 Euler / CG run "
                   "as ONE cooperative launch with the field resident in the SMs' shared memory (hbm_frac is then the "
                   "HBM-equivalent of their LUP rate); the single 1024^2 operator application is launch-latency bound; "
                   "cfg3 fp64_issue_frac = GLUP/s x non-fusable fp64 operations per LUP / (148 SMs x 64 lanes x 1965 MHz), "
                   "the bound of the bit-exact Euler step (ncu: fp64 pipe 67 % active at 256^3)")
    return out


def secondary_strong(rank, world, dev, hbm, n=1024, iters=100):
    """Config 5: CG on a fixed n^3 grid slab-decomposed over the ranks (strong scaling)."""
    import torch
    import torch.distributed as dist

    from pyapes_b200.geometry import Box
    from pyapes_b200.parallel import SlabMesh
    from pyapes_b200.solver.fdm import FDM
    from pyapes_b200.solver.ops import Solver
    from pyapes_b200.variables import Field
    from pyapes_b200.variables.bcs import homogeneous_bcs

    mesh = SlabMesh(Box[0:1, 0:1, 0:1], None, [n, n, n], rank, world, dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rhs = torch.rand((1, *mesh.nx), generator=g, dtype=torch.float64, device=dev)
    cfg = {"method": "cg", "tol": 1e-30, "max_it": iters - 1, "report": False, "check_every": iters + (iters & 1)}

    def run():
        var = Field("p", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
        s = Solver({"fdm": dict(cfg)})
        s.set_eq(FDM().laplacian(1.0, var) == rhs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rep = s.solve()
        assert rep["itr"] == iters, rep

    run()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 2], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    glups = float(n) ** 3 * iters / (t.item() * 1e-3) / 1e9
    del rhs, mesh
    torch.cuda.empty_cache()
    return {"cfg5_cg_1024_strong_glups": round(glups, 2), "cfg5_cg_1024_strong_ms_per_iter": round(t.item() / iters, 4),
            "cfg5_cg_1024_strong_per_gpu_hbm_frac": round(glups / world * B_PER_LUP_CG / hbm, 4),
            "cfg5_note": f"{n}^3 fp64 CG, {iters} iterations, {world} slabs of {n // world} planes; max over ranks"}


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

    e2e_trace = {}

    def _run_e2e(k_steps):
        ev_in = [None, None]
        # per-step events (timing enabled): what the copies and the solve of each step cost and where the step waits
        tr = {"up0": [], "up1": [], "s0": [], "s1": [], "dn0": [], "dn1": []}

        def upload(k):
            with torch.cuda.stream(h2d_s):
                a = torch.cuda.Event(enable_timing=True)
                a.record(h2d_s)
                rhs_buf[k % 2][:, olo:ohi].copy_(rhs_h, non_blocking=True)
                ev_in[k % 2] = torch.cuda.Event(enable_timing=True)
                ev_in[k % 2].record(h2d_s)
                tr["up0"].append(a)
                tr["up1"].append(ev_in[k % 2])

        upload(0)
        for k in range(k_steps):
            torch.cuda.current_stream().wait_event(ev_in[k % 2])
            if k + 1 < k_steps:
                upload(k + 1)  # overlaps this step's solve
            s0 = torch.cuda.Event(enable_timing=True)
            s0.record()
            solver, var = new_solver(rhs_buf[k % 2])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                solver.solve()  # returns when the device has finished this solve
            done = torch.cuda.Event(enable_timing=True)
            done.record()
            sol = var()
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(done)
                d0 = torch.cuda.Event(enable_timing=True)
                d0.record(d2h_s)
                out_h.copy_(sol[:, olo:ohi], non_blocking=True)  # overlaps the next solve
                d1 = torch.cuda.Event(enable_timing=True)
                d1.record(d2h_s)
            sol.record_stream(d2h_s)  # the allocator may recycle it only after the download
            tr["s0"].append(s0)
            tr["s1"].append(done)
            tr["dn0"].append(d0)
            tr["dn1"].append(d1)
            del solver, var, sol
        torch.cuda.current_stream().wait_stream(d2h_s)
        torch.cuda.synchronize()
        med = lambda v: sorted(v)[len(v) // 2] if v else None
        e2e_trace.clear()
        e2e_trace.update({
            "upload_ms_median": med([a.elapsed_time(b) for a, b in zip(tr["up0"], tr["up1"])]),
            "solve_ms_median": med([a.elapsed_time(b) for a, b in zip(tr["s0"], tr["s1"])]),
            "download_ms_median": med([a.elapsed_time(b) for a, b in zip(tr["dn0"], tr["dn1"])]),
            # device time between the end of one solve and the start of the next (host work + waiting for the upload)
            "gap_ms_median": med([tr["s1"][i].elapsed_time(tr["s0"][i + 1]) for i in range(len(tr["s0"]) - 1)]),
            "note": "this rank's events; copies of step k+1 / k-1 run on their own streams during the solve of step k"})

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

    # --- outside the timed regions: parity at N > 1, config 5 -----------------------------------
    hbm, peak_src = peaks()
    parity, secondary = {}, {}
    if world > 1 and not args.no_secondary:
        try:
            pd = parity_dist(rank, world, dev)
            if pd:
                parity.update(pd)
        except Exception as e:  # noqa: BLE001
            parity["dist_ok"] = False
            parity["dist_error"] = f"{type(e).__name__}: {e}"[:300]
        barrier()
        try:
            del rhs_buf, rhs_d
            torch.cuda.empty_cache()
            secondary.update(secondary_strong(rank, world, dev, hbm))
        except Exception as e:  # noqa: BLE001
            secondary["cfg5_error"] = f"{type(e).__name__}: {e}"[:300]
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- kernel-level roofline: every launch bracketed by CUDA events on the launching stream --
    from pyapes_b200 import profile as P

    kt = P.cg_kernel_times(n, iters=20)
    cells = float(n) ** 3
    ach_b = B_PER_CELL_PHASE_B * cells / (kt["phaseB_ms"] * 1e-3) / 1e9
    ach_a = B_PER_CELL_PHASE_A * cells / (kt["phaseA_ms"] * 1e-3) / 1e9
    traffic, traffic_note = ncu_traffic(f"phaseB_{n}")

    if world == 1 and args.no_secondary:
        cpu_obj = None
    elif world == 1:
        # cpu_baseline (rank 0 at N = 1 only) doubles as the config-2-size parity oracle
        cpu = cpu_baseline(args.cpu_n, args.cpu_iters, keep=True)
        try:
            parity.update(parity_vs_cpu(cpu, args.cpu_n, args.cpu_iters, dev))
        except Exception as e:  # noqa: BLE001
            parity["ok"] = False
            parity["error"] = f"{type(e).__name__}: {e}"[:300]
        cpu.pop("solution", None)
        cpu_obj = {"value": cpu["value"], "unit": "GLUP/s", "cores": cpu["threads"], "kind": cpu["kind"],
                   "sample": cpu["sample"]}
        del rhs_buf, rhs_d
        torch.cuda.empty_cache()
        sampler3 = ClockSampler(local)
        sampler3.start()
        time.sleep(0.3)
        secondary.update(secondary_single(hbm, sampler3))
        sampler3.stop()
    else:
        cpu_obj = None  # (the CPU arm is `--impl reference`; at N > 1 it is one host process on rank 0)

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
                     "frac": ach_b / hbm, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                     "avg_launch_ms": kt["phaseB_ms"], "algorithmic_bytes_per_launch": B_PER_CELL_PHASE_B * cells,
                     "phaseA_avg_launch_ms": kt["phaseA_ms"], "phaseA_achieved": ach_a, "phaseA_frac": ach_a / hbm,
                     "share_phaseA": kt["share"]["phaseA"], "share_phaseB": kt["share"]["phaseB"],
                     "other_kernels": {kt["kernels"][0]: {"avg_launch_ms": kt["phaseA_ms"], "achieved": ach_a,
                                                               "frac": ach_a / hbm},
                                       "bc_faces+shell_norm_ms_per_iter": kt["small_ms"]},
                     "kernel_share_of_iteration": kt["share"]},
        "e2e": {"value": e2e_value, "unit": "GLUP/s", "h2d_bytes_per_step": int(rhs_h.numel() * 8),
                "d2h_bytes_per_step": int(out_h.numel() * 8), "clocks": clocks_e2e, "trace": dict(e2e_trace),
                "h2d_GBps": copy_rates["h2d_GBps"], "d2h_GBps": copy_rates["d2h_GBps"],
                "pinned_copy_rates": copy_rates,
                "note": "copies of step k+1 / k-1 overlap the solve of step k; the first upload and the last "
                        "download are exposed"},
        "gpu_launches": int(launches_timed),
        "clocks": clocks,
        "secondary": secondary,
        "parity": parity,
    }
    if cpu_obj is not None:
        out["cpu_baseline"] = cpu_obj
    else:
        out["cpu_baseline"] = {"value": None, "unit": "GLUP/s", "cores": 0, "kind": "see --impl reference",
                               "sample": "not run (rank 0 at N = 1 only; skipped by --no-secondary)"}
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
    ap.add_argument("--no-secondary", action="store_true",
                    help="headline only: skip `secondary`, `parity` and the CPU baseline (profiling runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
