"""Static instruction mix of the LEAN and the general (edge-tile) path of the fused TMA kernels, read
from the in-tree library with cuobjdump -- no GPU needed.  The two paths are consecutive runs of fp64
multiplies in the SASS; the kernel's time at 512^3 followed 0.7 * LEAN + 0.3 * general to 0.2 %
when the general path lost its periodic-wrap code, so this is the first thing to
look at after touching kernels_tma.cuh.
usage: python tools/sass_paths.py ['regex on the demangled kernel name']"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyapes_b200", "lib", "libpyapes_b200.so")
KEYS = ("DMUL", "DADD", "FSEL", "ISETP", "IMAD", "LDC", "LDS", "LDG", "STG", "LDL", "STL", "BRA")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_.]+)", line)
        if cur and m:
            funcs[cur].append(m.group(2).split(".")[0])
    names = list(funcs)
    plain = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return {p.split("(")[0].replace("void pa::", ""): funcs[n] for p, n in zip(plain, names)}


def paths(ins):
    """Split the first dense run of DMULs into its two equal halves (LEAN, general)."""
    dm = [k for k, op in enumerate(ins) if op == "DMUL"]
    if len(dm) < 20:
        return []
    end = len(dm)
    for j in range(1, len(dm)):
        if dm[j] - dm[j - 1] > 300:
            end = j
            break
    dm = dm[:end]
    half = len(dm) // 2
    out = []
    for a, b in ((dm[0], dm[half - 1]), (dm[half], dm[2 * half - 1])):
        c = Counter(ins[a:b + 1])
        out.append({"instructions": b - a + 1, **{k: c[k] for k in KEYS}})
    return out


if __name__ == "__main__":
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_cg_phase[AB]_tma<double, pa::KStd")
    for name, ins in sorted(functions().items()):
        if pat.search(name):
            print(name)
            for label, row in zip(("LEAN   ", "general"), paths(ins)):
                print("   ", label, row)
