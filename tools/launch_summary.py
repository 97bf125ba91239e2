"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: one device-resident forward, delimited by
two consecutive pack kernels, grouped by kernel.  Usage: python -m tools.launch_summary gpurun_out/launches.csv [nth forward]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
nth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((r["Kernel Name"], us))
packs = [i for i, (k, _) in enumerate(rows) if "pack_" in k]
a, b = packs[nth - 1], packs[nth]
sel = rows[a:b]
agg = collections.OrderedDict()
for k, us in sel:
    k = re.sub(r"^void (vrd::)?(<unnamed>::)?", "", k)
    k = re.sub(r"\(.*$", "", k)[:64]
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += us
tot = sum(us for _, us in sel)
print(f"launches {a}..{b} of {len(rows)} ({len(sel)} launches, {tot:.1f} us; cold-cache serialised launches under ncu: compare SHARES)")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} {n:4d} launches {us:9.1f} us {100 * us / tot:5.1f}%")
