"""Anomaly hunt: CG 1024^2 resident, repeated, with and without the per-iteration x stores."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

def cg(n, variant, tag):
    r = P.solver_throughput([n, n], "cg", 1000, variant=variant)
    print(json.dumps({"case": f"cg {n}^2 v{variant} {tag}", "us_per_iter": round(r["ms"], 3)}), flush=True)

for rep in range(3):
    cg(1024, 6, "plain")
cg(512, 4, "")
cg(1024, 6, "after 512 v4")
cg(1024, 4, "")
cg(1024, 6, "after 1024 v4")
os.environ["PA_RES_DEBUG_FLAGS"] = "1"
cg(1024, 6, "no x stores")
cg(1024, 6, "no x stores")
os.environ.pop("PA_RES_DEBUG_FLAGS")
os.environ["PA_RES_DEBUG"] = "1"
cg(1024, 6, "debug stamps")
