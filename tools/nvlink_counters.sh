#!/bin/bash
# NVLink data counters of every GPU around a command: evidence that the halo planes and the all-reduce
# mailboxes travel over NVLink (no NCCL call in the CG loop).  usage: tools/nvlink_counters.sh <out> <cmd...>
out=$1; shift
nvidia-smi nvlink -gt d > "$out.before" 2>&1
"$@"
rc=$?
nvidia-smi nvlink -gt d > "$out.after" 2>&1
python - "$out" <<'PY'
import re, sys
out = sys.argv[1]
def parse(path):
    gpu, tot = None, {}
    for line in open(path):
        m = re.match(r"GPU (\d+):", line)
        if m: gpu = int(m.group(1)); tot.setdefault(gpu, [0, 0]); continue
        m = re.search(r"Link \d+: Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None: tot[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return tot
b, a = parse(out + ".before"), parse(out + ".after")
with open(out + ".txt", "w") as f:
    for g in sorted(a):
        tx, rx = a[g][0] - b.get(g, [0, 0])[0], a[g][1] - b.get(g, [0, 0])[1]
        line = f"GPU {g}: NVLink Tx {tx / 2**20:.3f} GiB  Rx {rx / 2**20:.3f} GiB during the command"
        print(line); f.write(line + "\n")
PY
exit $rc
