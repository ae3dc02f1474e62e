"""Anomaly hunt 2: does a cooperative TMA / persistent small-grid CG launch earlier in the process slow the resident CG?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402


def cg(n, variant, tag=""):
    r = P.solver_throughput([n, n], "cg", 1000, variant=variant)
    print(json.dumps({"case": f"cg {n}^2 v{variant} {tag}", "us_per_iter": round(r["ms"], 3)}), file=sys.stderr, flush=True)


cg(1024, 6, "first")
cg(512, 5)
cg(1024, 6, "after 512 v5")
cg(256, 3)
cg(1024, 6, "after 256 v3")
cg(512, 4)
cg(1024, 6, "after 512 v4")
for n in (256, 512):
    for v in (6, 5, 4):
        cg(n, v)
os.environ["PA_RES_DEBUG"] = "1"
cg(1024, 6, "after the bench_resident sequence (debug stamps)")
