"""Write profiles/ncu_traffic.json from an `ncu --set full` capture of the headline kernels: per-launch
dram__bytes_read.sum + dram__bytes_write.sum of CG phase B / phase A, stamped with the content hash of the
sources those kernels are compiled from (__graft_entry__._cg_kernel_hash: kernels_tma.cuh, what it includes, the two
translation units, the nvcc flags); bench.py refuses a figure whose hash is not the current one.
usage: python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep <n> [label]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G

rep, n = sys.argv[1], int(sys.argv[2])
label = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


acc = {}
for r in rows[2:]:
    for tag, pat in (("phaseB", "k_cg_phaseB_tma"), ("phaseA", "k_cg_phaseA_tma")):
        if pat in r[ik]:
            acc.setdefault(tag, []).append((to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]), float(r[it].replace(",", ""))))
out = {"source_hash": G._cg_kernel_hash(), "capture": label, "how": "ncu --set full --clock-control none; mean over the captured launches"}
for tag, v in acc.items():
    out[f"{tag}_{n}"] = int(sum(b for b, _ in v) / len(v))
    out[f"{tag}_{n}_launches"] = len(v)
    out[f"{tag}_{n}_ncu_duration_{units[it]}"] = round(sum(t for _, t in v) / len(v), 2)
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
