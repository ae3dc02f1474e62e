import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyapes_b200 import profile as P
print(P.cg_kernel_times(512, iters=3, dtype="single"))
