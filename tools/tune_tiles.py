"""Run the instrumented CG pass for several (library variant, PA_TILE_RY) combinations."""
import os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "512"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[2:] or ["libpyapes_b200.so"]:
    for ry in ("2", "4"):
        env = dict(os.environ, PA_LIB=os.path.join(root, "pyapes_b200", "lib", lib), PA_TILE_RY=ry)
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "prof_cg.py"), n, "10", "nobuild"],
                             env=env, capture_output=True, text=True)
        print(lib, "RY", ry, (out.stdout.strip().splitlines() or [out.stderr[-400:]])[-1])
