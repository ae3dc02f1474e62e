"""Run-to-run spread of the fp32 CG 512^3 solve (bench.py secondary `cg_512_fp32` has shown 108-183 GLUP/s)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

for rep in range(6):
    r = P.solver_throughput([512] * 3, "cg", 200, dtype="single")
    print(json.dumps({"case": "cg 512^3 fp32 200 it", "ms": round(r["ms"], 2), "GLUP/s": round(r["GLUP/s"], 1)}), flush=True)
r = P.solver_throughput([512] * 3, "cg", 200, dtype="single", reps=5)
print(json.dumps({"case": "cg 512^3 fp32 200 it, mean of 5", "ms": round(r["ms"], 2), "GLUP/s": round(r["GLUP/s"], 1)}), flush=True)
k = P.cg_kernel_times(512, iters=20, dtype="single")
print(json.dumps({k2: (round(v, 4) if isinstance(v, float) else v) for k2, v in k.items() if k2 not in ("share", "kernels")}), flush=True)
