"""Summarise an ncu --csv launch list (gpu__time_duration.sum [, smsp__inst_executed.sum]) per kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
t = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    rec = dict(zip(H, r))
    name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "").replace("pf::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    v = float(rec["Metric Value"].replace(",", ""))
    if "gpu__time_duration" in rec["Metric Name"]:
        u = rec["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        t[name][0] += 1
        t[name][1] += v
        if v > 8.0:
            t[name][3] += v
    elif "inst_executed" in rec["Metric Name"]:
        t[name][2] += v
tot = sum(v[1] for v in t.values())
print("%-34s %6s %10s %6s %12s %10s" % ("kernel", "n", "us", "%", "Mwarp-instr", "us in >8us launches"))
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %6d %10.1f %5.1f%% %12.2f %10.1f" % (k[:34], v[0], v[1], 100 * v[1] / tot, v[2] / 1e6, v[3]))
print("total %.1f us; without sweeps %.1f us" % (tot, sum(v[1] for k, v in t.items() if "sweep6" not in k)))
