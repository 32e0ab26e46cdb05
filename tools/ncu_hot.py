#!/usr/bin/env python3
"""Top stall hot-spots of an ncu source-page CSV: usage ncu_hot.py <source.csv> [topN]
(ncu -i rep --page source --csv > source.csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
for i, r in enumerate(body):            # several kernels in one report: keep the first section only
    if r and r[0] == "Kernel Name":
        body = body[:i]
        break
body = [r for r in body if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
execd = {}
for r in body:
    op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"
    if op.startswith("@"): op = r[ci["Source"]].split()[1]
    execd[op.split(".")[0]] = execd.get(op.split(".")[0], 0) + int(r[ci["Instructions Executed"]] or 0)
te = sum(execd.values())
print("executed warp-instrs:", te, {k: round(v / te, 4) for k, v in sorted(execd.items(), key=lambda kv: -kv[1])[:14]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    why = {s[6:]: int(r[ci[s]]) for s in stalls if int(r[ci[s]] or 0) > 0}
    print("%5d %6s %5.1f%%  %-70s %s" % (i, r[ci["# Samples"]], 100.0 * int(r[ci["# Samples"]]) / tot, r[ci["Source"]].strip()[:70], why))
