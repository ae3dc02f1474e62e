"""A few explicit operator applications for ncu. usage: python tools/prof_apply.py [op] [n] [dtype] [reps]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from pyapes_b200 import profile as P
op = sys.argv[1] if len(sys.argv) > 1 else "laplacian"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dt = sys.argv[3] if len(sys.argv) > 3 else "double"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
shape = [n] * 3 if n <= 1024 else [n, n]
print(P.operator_apply_times(shape, op, dt, reps=reps))
