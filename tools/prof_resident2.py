"""Long resident runs with in-kernel time stamps against the event time (anomaly hunt)."""
import os
import sys
import json

os.environ["PA_RES_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyapes_b200.profile as P  # noqa: E402

for rep in range(3):
    for n in (512, 1024):
        r = P.solver_throughput([n, n], "cg", 1000, variant=6)
        print(json.dumps({"case": f"cg {n}^2 v6", "us_per_iter_events": round(r["ms"], 3)}), file=sys.stderr, flush=True)
    r = P.euler_throughput([1024, 1024], "upwind", 2000)
    print(json.dumps({"case": "euler 1024^2", "us_per_step_events": round(r["ms"] / 2, 3)}), file=sys.stderr, flush=True)
