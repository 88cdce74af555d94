"""Summarise an ncu --csv launch list per kernel: launches, time, warp instructions, DRAM bytes.

    python tools/summarise_launches.py profiles/r2_launches_4000x2000.csv [--sweep-traffic profiles/sweep_traffic.json]

--sweep-traffic writes the average DRAM bytes per launch of the wavefront sweep kernel (dram__bytes_read.sum + dram__bytes_write.sum),
which bench.py reports as roofline.traffic."""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
t = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])      # launches, us, warp instructions, us in launches > 8 us, DRAM bytes
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    rec = dict(zip(H, r))
    name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "").replace("pf::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    v = float(rec["Metric Value"].replace(",", ""))
    m, u = rec["Metric Name"], rec["Metric Unit"]
    if "gpu__time_duration" in m:
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        t[name][0] += 1
        t[name][1] += v
        if v > 8.0:
            t[name][3] += v
    elif "inst_executed" in m:
        t[name][2] += v
    elif "dram__bytes" in m:
        t[name][4] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
tot = sum(v[1] for v in t.values())
print("%-34s %6s %10s %6s %12s %12s %10s" % ("kernel", "n", "us", "%", "Mwarp-instr", "DRAM MB", "us in >8us launches"))
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %6d %10.1f %5.1f%% %12.2f %12.1f %10.1f" % (k[:34], v[0], v[1], 100 * v[1] / tot, v[2] / 1e6, v[4] / 1e6, v[3]))
is_sweep = lambda k: k.startswith("k_sweep<") or k.startswith("k_sweep6") or k.startswith("k_sweep7")
print("total %.1f us, %.1f Mwarp-instr, %.1f MB DRAM; without sweeps %.1f us" % (
    tot, sum(v[2] for v in t.values()) / 1e6, sum(v[4] for v in t.values()) / 1e6, sum(v[1] for k, v in t.items() if not is_sweep(k))))
if "--sweep-traffic" in sys.argv:
    out = sys.argv[sys.argv.index("--sweep-traffic") + 1]
    n = sum(v[0] for k, v in t.items() if is_sweep(k))
    b = sum(v[4] for k, v in t.items() if is_sweep(k))
    us = sum(v[1] for k, v in t.items() if is_sweep(k))
    json.dump({"dram_bytes_per_launch": b / max(1, n), "launches": n, "avg_launch_us_isolated": us / max(1, n),
               "source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the %d sweep launches of one 4000x2000 pair (ncu launch list %s)" % (n, sys.argv[1])},
              open(out, "w"), indent=1)
