"""Per-kernel CG times (CUDA events around every launch) for fp64 / fp32, bit-exact / contraction mode.
usage: python tools/prof_variants.py [n]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from pyapes_b200 import profile as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for dtype in ("double", "single"):
    for variant, tag in ((0, "exact"), (0x100, "contract")):
        r = P.cg_kernel_times(n, iters=20, variant=variant, dtype=dtype)
        esz = 8 if dtype == "double" else 4
        cells = float(n) ** 3
        print(dtype, tag, "A %.4f ms (%.3f) B %.4f ms (%.3f) iter %.4f ms -> %.1f GLUP/s" % (
            r["phaseA_ms"], 3 * esz * cells / r["phaseA_ms"] / 1e6 / 6541.8, r["phaseB_ms"], 5 * esz * cells / r["phaseB_ms"] / 1e6 / 6541.8,
            r["iter_ms"], cells / r["iter_ms"] / 1e6), flush=True)
    torch.set_default_dtype(torch.float64)
