"""Run-to-run stability of the headline kernels inside one process: per-kernel CG times, 6 alternating passes.
usage: python tools/prof_stability.py [n]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from pyapes_b200 import profile as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for k in range(6):
    for variant, tag in ((0, "exact"), (0x100, "contract")):
        r = P.cg_kernel_times(n, iters=20, variant=variant)
        print(k, tag, "A %.4f B %.4f iter %.4f ms" % (r["phaseA_ms"], r["phaseB_ms"], r["iter_ms"]), flush=True)
for contract in (False, True, False, True):
    r = P.solver_throughput([n] * 3, "cg", 200, reps=3, contract=contract)
    print("solve x3 contract=%s: %.2f GLUP/s" % (contract, r["GLUP/s"]), flush=True)
