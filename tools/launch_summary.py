"""per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> markdown (shares of the serialised time).
   python tools/launch_summary.py launches.csv [title] > summary.md"""
import collections, csv, io, re, sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else ""
lines = open(path, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("cruse::", "").replace("(anonymous namespace)::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name] += us
    cnt[name] += 1
allus = sum(tot.values())
print(f"ncu launch list {title}(`{path}`; cold-cache, serialised launches: compare SHARES, not absolutes)\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in tot.most_common():
    print(f"| {k} | {cnt[k]} | {v:.1f} | {100 * v / allus:.1f}% |")
print(f"\n{sum(cnt.values())} launches, {allus:.1f} us")
