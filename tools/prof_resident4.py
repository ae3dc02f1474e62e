"""Anomaly hunt: CG 1024^2 resident with phase stamps, then variants (no x stores, item loop, 1000x1024)."""
import json
import os
import sys

os.environ["PA_RES_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402


def cg(shape, tag, iters=1000):
    r = P.solver_throughput(shape, "cg", iters, variant=6)
    print(json.dumps({"case": f"cg {shape} v6 {tag}", "us_per_iter": round(r["ms"] * 1e3 / iters, 3)}), file=sys.stderr, flush=True)
    return r["ms"] * 1e3 / iters


t = cg([1024, 1024], "plain")
cg([512, 512], "plain")
cg([1000, 1024], "R=7, 143 CTAs")
cg([888, 1024], "R=6, 148 CTAs")
cg([1024, 512], "R=7 half rows")
os.environ["PA_RES_DEBUG_FLAGS"] = "1"
cg([1024, 1024], "no x stores")
os.environ.pop("PA_RES_DEBUG_FLAGS")
os.environ["PA_RES_PATH"] = "items"
cg([1024, 1024], "item loop")
os.environ.pop("PA_RES_PATH")
cg([1024, 1024], "plain again")
r = P.solver_throughput([1024, 1024], "jacobi", 1000, variant=6)
print(json.dumps({"case": "jacobi 1024^2 v6", "us_per_sweep": round(r["ms"], 3)}), file=sys.stderr, flush=True)
print("ANOMALY" if t > 40 else "normal", file=sys.stderr)
