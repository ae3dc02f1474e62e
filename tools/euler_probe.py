"""Why do two Euler runs of the SAME kernel (k_star_tma<PW_EULER,2>) differ in speed?  Times the
explicit adv-diff step at one size for several coefficient sets while sampling SM clocks / power.
usage: python tools/euler_probe.py [n]"""
import os, sys, threading, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.filterwarnings("ignore")
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import homogeneous_bcs
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = [n] * 3


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.stop = False; self.clk = []; self.pw = []
    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0)
            time.sleep(0.005)


def case(limiter, u, steps=200, init="rand"):
    mesh = Mesh(Box([0.0] * 3, [1.0] * 3), None, shape, "cuda")
    var = Field("c", 1, mesh, {"domain": homogeneous_bcs(3, 0.0, "dirichlet"), "obstacle": None})
    g = torch.Generator().manual_seed(1234)
    x0 = torch.rand(1, *shape, generator=g, dtype=torch.float64).cuda()
    if init == "zero":
        x0.zero_()
    var.set_var_tensor(x0)
    nu = 0.1
    var.set_time(0.2 * min(mesh._dx) ** 2 / nu, 0.0)
    fdm = FDM({"div": {"limiter": limiter, "edge": False}})
    s = Solver({"fdm": {"method": "euler", "tol": 0.0, "max_it": 0, "report": False, "n_steps": steps}})
    s.set_eq(fdm.ddt(var) + fdm.div(u, var) - fdm.laplacian(nu, var) == 0.0)
    s.solve(); torch.cuda.synchronize()
    smp = Sampler(); smp.start()
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    smp.stop = True; smp.join()
    v = var()
    clk = sorted(smp.clk)[len(smp.clk) // 2] if smp.clk else -1
    print(f"{limiter:10s} u={u:+.1f} init={init:5s} {best / steps * 1e3:8.1f} us/step  {n ** 3 * steps / best / 1e6:7.1f} GLUP/s  "
          f"sm_clk~{clk} MHz  power~{max(smp.pw) if smp.pw else -1:.0f} W  finite={bool(torch.isfinite(v).all())} "
          f"absmax={v.abs().max().item():.3e} subnormal_frac={(v.abs() < 2.3e-308).logical_and(v != 0).double().mean().item():.2e}", flush=True)


case("upwind", 1.0)
case("upwind_fd", 1.0)
case("upwind", -1.0)
case("upwind_fd", -1.0)
case("upwind_fd", 1.0, init="zero")
case("none", 1.0)
