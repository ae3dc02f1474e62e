"""Anomaly hunt 3: the bench_resident.py prefix (Euler 256^2 first), then CG 1024^2; diagnostics in the SAME process
when the anomaly shows (> 40 us per iteration)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402


def say(**kw):
    print(json.dumps(kw), file=sys.stderr, flush=True)


def cg(shape, tag, iters=1000, variant=6):
    r = P.solver_throughput(shape, "cg", iters, variant=variant)
    say(case=f"cg {shape} v{variant} {tag}", us_per_iter=round(r["ms"] * 1e3 / iters, 3))
    return r["ms"] * 1e3 / iters


r = P.euler_throughput([256, 256], "upwind", 2000)
say(case="euler 256^2 first", us_per_step=round(r["ms"] / 2, 3))
t = cg([1024, 1024], "plain")
if t < 40:
    say(verdict="normal")
    sys.exit(0)
say(verdict="ANOMALY")
os.environ["PA_RES_DEBUG"] = "1"
cg([1024, 1024], "stamps", iters=1000)
os.environ.pop("PA_RES_DEBUG")
cg([1024, 1024], "again")
cg([1000, 1024], "143 CTAs")
cg([888, 1024], "R=6, 148 CTAs")
cg([1024, 512], "half rows")
cg([512, 512], "512^2")
os.environ["PA_RES_PATH"] = "items"
cg([1024, 1024], "item loop")
os.environ.pop("PA_RES_PATH")
r = P.solver_throughput([1024, 1024], "jacobi", 1000, variant=6)
say(case="jacobi 1024^2 v6", us_per_sweep=round(r["ms"], 3))
r = P.euler_throughput([1024, 1024], "upwind", 2000)
say(case="euler 1024^2", us_per_step=round(r["ms"] / 2, 3))
cg([1024, 1024], "last")
