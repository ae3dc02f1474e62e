"""Short CG run for ncu: n^3 Dirichlet Poisson, a few iterations of the tiled kernels.
usage: python tools/prof_cg.py [n] [iters]"""
import sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as G
if 'nobuild' not in sys.argv:
    G.build()
from pyapes_b200 import profile as P
args = [a for a in sys.argv[1:] if a != "nobuild"]
n = [int(v) for v in args[0].split("x")] if args else [256]
n = n[0] if len(n) == 1 else n
iters = int(args[1]) if len(args) > 1 else 3
r = P.cg_kernel_times(n, iters=iters)
cells = n ** 3 if isinstance(n, int) else n[0] * n[1] * n[2]
r["GLUP/s"] = cells / (r["iter_ms"] * 1e-3) / 1e9
print(n, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("share", "kernels")})
