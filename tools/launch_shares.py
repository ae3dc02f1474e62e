"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python tools/launch_shares.py launches.csv "<command line profiled>" > profiles/....txt"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        name = r[ki].split("(")[0][:60]
        tot[name] += float(r[vi].replace(",", "")) / 1e3
        cnt[name] += 1
total = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
for name, t in tot.most_common():
    print(f"{name:60s} {cnt[name]:8d} {t:12.1f} {100 * t / total:6.1f}% {t / cnt[name]:10.1f}")
