"""Short CG run for ncu: n^3 Dirichlet Poisson, a few iterations of the tiled kernels.
usage: python tools/prof_cg.py [n] [iters]"""
import sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as G
if 'nobuild' not in sys.argv:
    G.build()
from pyapes_b200 import profile as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
print(P.cg_kernel_times(n, iters=iters))
