"""Resident (shared-memory) kernels against the streaming paths on 2-D grids: Euler steps and CG iterations.
   python tools/bench_resident.py  -> one JSON line per case"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

for n in (256, 512, 1024):
    for v in ("resident", "stream"):
        if v == "stream":
            os.environ["PA_EULER_VARIANT"] = "stream"
        else:
            os.environ.pop("PA_EULER_VARIANT", None)
        r = P.euler_throughput([n, n], "upwind", 2000)
        print(json.dumps({"case": f"euler {n}^2 {v}", "GLUP/s": round(r["GLUP/s"], 2), "us_per_step": round(r["ms"] * 1e3 / 2000, 3)}), flush=True)
os.environ.pop("PA_EULER_VARIANT", None)
for n in (256, 512, 1024):
    for variant in (6, 5, 4, 3):
        if variant == 3 and n > 256:
            continue
        r = P.solver_throughput([n, n], "cg", 1000, variant=variant)
        print(json.dumps({"case": f"cg {n}^2 variant {variant}", "GLUP/s": round(r["GLUP/s"], 2), "us_per_iter": round(r["ms"] * 1e3 / 1000, 3),
                          "launches": r["launches"]}), flush=True)
