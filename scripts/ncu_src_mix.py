#!/usr/bin/env python
"""Instruction mix and stall samples of one kernel from the SASS source page of an .ncu-rep:
    ncu -i rep --page source --csv --print-source sass > src.csv ; ncu_src_mix.py src.csv
Prints executed warp instructions by class (FP64 / LDS / SHFL / local / other), and the same per "region": maximal runs of
consecutive instructions with (nearly) the same execution count, i.e. loop bodies, so that the cost of the term loop, the
row step and the block prologue / epilogue can be read separately."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {k: i for i, k in enumerate(rows[hdr])}
ins = []
for r in rows[hdr + 1:]:
    if len(r) < len(col): continue
    text = r[col["Source"]].strip()
    op = text.split()[1] if text.startswith("@") else text.split()[0]
    ins.append(dict(text=text, op=op.split(".")[0], n=int(r[col["Instructions Executed"]]), samples=int(r[col["# Samples"]]),
                    wait=int(r[col["stall_wait"]]), math=int(r[col["stall_math"]]), short=int(r[col["stall_short_sb"]]),
                    notsel=int(r[col["stall_not_selected"]]), sel=int(r[col["stall_selected"]]), mio=int(r[col["stall_mio"]])))
def cls(op):
    if op in ("DFMA", "DMUL", "DADD"): return "fp64"
    if op == "LDS": return "lds"
    if op == "SHFL": return "shfl"
    if op in ("LDL", "STL"): return "local"
    return "other"
tot = sum(x["n"] for x in ins); ts = sum(x["samples"] for x in ins)
mix = {}
for x in ins: mix[cls(x["op"])] = mix.get(cls(x["op"]), 0) + x["n"]
print("executed warp instructions: %.4e" % tot, {k: "%.1f%%" % (100 * v / tot) for k, v in mix.items()})
print("stall samples: total %d" % ts, {k: "%.1f%%" % (100 * sum(x[k] for x in ins) / ts) for k in ("wait", "math", "notsel", "sel", "short", "mio")})
# regions by execution count
regions, cur = [], None
for x in ins:
    if cur and x["n"] > 0 and abs(x["n"] - cur["n0"]) <= 0.02 * cur["n0"]:
        cur["ins"].append(x)
    else:
        cur = dict(n0=max(x["n"], 1), ins=[x]); regions.append(cur)
print("regions (>= 0.5 %% of the instructions):")
for g in regions:
    n = sum(x["n"] for x in g["ins"])
    if n < 0.005 * tot: continue
    m = {}
    for x in g["ins"]: m[cls(x["op"])] = m.get(cls(x["op"]), 0) + 1
    s = sum(x["samples"] for x in g["ins"])
    print("  exec/instr %.3e  len %4d  share of instr %5.1f%%  of samples %5.1f%%  %s  first: %s" % (
        g["n0"], len(g["ins"]), 100 * n / tot, 100 * s / ts, m, g["ins"][0]["text"][:50]))
