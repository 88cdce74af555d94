"""One unrolled wavefront step of k_sweep<1,1> from an ncu source-page export, every instruction classified.

    ncu -i gpurun_out/r2f_sweep_v12.ncu-rep --page source --csv > /tmp/src.csv
    python tools/classify_hot_loop.py /tmp/src.csv "<header text>" > profiles/r2_sweep_hot_loop.sass

The step is delimited by the skip vote (first VOTE.ANY of a step) of two consecutive unrolled steps.  Classes are assigned from the
opcode (and, for the few ambiguous ones, the operands): what the step's dependent chain needs (gather, arithmetic, exchange, select)
against what only keeps the pipeline around it running (addresses, range checks of the exact sequences, control flow, moves,
records / rings / stores)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
ix = {h: i for i, h in enumerate(H)}
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[2:]:
    if len(r) < len(H):
        continue
    ins.append({"addr": int(r[0], 16), "sass": " ".join(r[1].split()), "samples": int(r[ix["# Samples"]]),
                "exec": int(r[ix["Instructions Executed"]]), "st": {s[6:]: int(r[ix[s]]) for s in stalls}})
base = ins[0]["addr"]
votes = [i for i, d in enumerate(ins) if "VOTE.ANY" in d["sass"]]
a, b = votes[0], votes[2]
step = ins[a:b]
hot = max(d["exec"] for d in step)


def classify(s):
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
    o = op.split(".")[0]
    if o in ("BRA", "BSSY", "BSYNC", "YIELD", "CALL", "RET", "EXIT", "WARPSYNC", "NANOSLEEP", "BMOV"):
        return "bookkeeping: control flow"
    if o == "UMOV":
        return "bookkeeping: shuffle/vote glue"
    if o == "VOTE":
        return "range check / skip vote"
    if o == "SHFL":
        return "chain: exchange"
    if o == "LDG":
        return "chain: gather"
    if o in ("LDS", "STS", "ST", "STG", "LDGSTS", "UBLKCP", "SYNCS", "LDC", "ULDC", "LD"):
        return "bookkeeping: records / rings / stores / L1 warm-up"
    if o in ("MOV",) or op.startswith("IMAD.MOV"):
        return "register moves"
    if o in ("VIADDMNMX", "VIMNMX", "VIMNMX3", "FMNMX3") or (o == "VIADD" and "0xffffffff" in s) or (o == "IADD3" and "-0x1" in s):
        return "range check of the exact sequences"
    if o in ("FFMA2", "FADD2", "FFMA", "FADD", "FMUL", "MUFU"):
        return "chain: arithmetic"
    if o in ("FMNMX", "F2I", "FRND", "I2F"):
        return "chain: bilinear cell (clamp, floor, fraction)"
    if o in ("FSEL", "FSETP", "SEL", "PLOP3"):
        return "chain: select / compare"
    if o in ("IMAD", "IADD3", "LEA", "LOP3", "SHF", "ISETP", "VIADD", "IABS", "PRMT"):
        return "address / index arithmetic"
    return "other"


tot_samples = sum(d["samples"] for d in step)
agg = {}
for d in step:
    d["cls"] = classify(d["sass"])
    if d["exec"] >= 0.9 * hot:
        e = agg.setdefault(d["cls"], [0, 0])
        e[0] += 1
        e[1] += d["samples"]
    else:
        e = agg.setdefault("(rarely executed: spin iterations, IEEE-intrinsic redo)", [0, 0])
        e[1] += d["samples"]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else "hot loop of k_sweep<1,1>"))
print("#")
print("# Columns: offset, SASS, class, times executed, stall samples, the two largest stall reasons of the instruction.")
print("# \"exec\" well below %d marks code that is skipped on most steps (spin iterations, the IEEE-intrinsic redo, back-pressure)." % hot)
print("#")
print("# Per class (instructions executed on every step / share of the step's stall samples):")
n_all = 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("#   %-62s %4d instructions  %5.1f %% of samples" % (k, v[0], 100.0 * v[1] / max(1, tot_samples)))
    n_all += v[0]
chain = sum(v[0] for k, v in agg.items() if k.startswith("chain"))
print("#   TOTAL: %d instructions executed on (nearly) every step, %d of them the step's own chain (cell, gather, arithmetic, exchange, select); "
      "%d static instructions, %d stall samples" % (n_all, chain, len(step), tot_samples))
print("#")
for d in step:
    top = sorted(d["st"].items(), key=lambda kv: -kv[1])[:2]
    print("%05x  %-72s %-58s exec=%-7d samples=%-4d %s" % (d["addr"] - base, d["sass"][:72], d["cls"], d["exec"], d["samples"],
                                                            " ".join("%s=%d" % (k, v) for k, v in top if v)))
