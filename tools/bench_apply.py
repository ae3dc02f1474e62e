"""Throughput of the explicit operator applications (Solver.Aop, FDC().laplacian/.grad/.div) against
their HBM roofline (2 words per cell, Grad 1 + d), TMA star engine vs generic kernels.
usage: python tools/bench_apply.py [--quick]"""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from pyapes_b200 import profile as P

HBM = 6541.8
if os.path.exists("MEASURED_PEAKS.json"):
    HBM = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
quick = "--quick" in sys.argv
MIXED = (["periodic", "periodic", "neumann", "symmetry", "dirichlet", "dirichlet"], [None, None, 0.5, None, 0.0, 0.0])
shapes = [([512] * 3, "double"), ([256] * 3, "double"), ([1024, 1024], "double"), ([4096, 4096], "double"),
          ([512] * 3, "single")]
if quick:
    shapes = shapes[1:3]
for shape, dt in shapes:
    for op in ("laplacian", "grad", "div_upwind", "advdiff"):
        for variant in (("auto", "tma", "generic") if len(shape) == 2 else ("tma", "generic")):
            r = P.operator_apply_times(shape, op, dt, reps=10 if len(shape) == 3 and shape[0] >= 512 else 30,
                                       variant=variant)
            r["hbm_frac"] = round(r["GB/s"] / HBM, 3)
            r["ms"] = round(r["ms"], 4); r["GLUP/s"] = round(r["GLUP/s"], 1); r["GB/s"] = round(r["GB/s"], 1)
            print(json.dumps(r), flush=True)
    torch.set_default_dtype(torch.float64)
# config-4 boundary conditions (periodic x, Neumann / Symmetry y): the general path of the edge tiles
r = P.operator_apply_times([512] * 3, "laplacian", "double", reps=10, kinds=MIXED[0], vals=MIXED[1])
r["hbm_frac"] = round(r["GB/s"] / HBM, 3); r["bcs"] = "config 4 mixed"
print(json.dumps(r), flush=True)
