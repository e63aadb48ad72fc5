#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source sass` export: executed warp instructions, stall samples and
shared-memory wavefronts per region of the kernel, regions being cut wherever the executed count changes by more than
a factor (so loops and phases show up as rows).  Usage: ncu_sass_regions.py export.csv [frames]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
ix = {k: i for i, k in enumerate(h)}
data = rows[hdr_i + 1:]
def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0
tot_exec = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print(f"total executed warp instructions {tot_exec:.4g}; samples {tot_samp:.0f}" + (f"; per frame {tot_exec / frames:.2f} warp / {tot_exec * 32 / frames:.0f} thread" if frames else ""))
# regions
regions = []
cur = None
for n, r in enumerate(data):
    e = f(r, "Instructions Executed")
    if cur is None or not (0.7 * cur["e0"] <= e <= 1.4 * cur["e0"]) :
        cur = {"start": n, "e0": max(e, 1.0), "rows": []}
        regions.append(cur)
    cur["rows"].append(r)
print(f"{'sass#':>6} {'n':>5} {'exec/instr':>11} {'exec':>10} {'%exec':>6} {'%samp':>6} {'wavefr':>10} {'excess':>10}  top opcodes / top stalls")
stall_keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for g in regions:
    ex = sum(f(r, "Instructions Executed") for r in g["rows"])
    sm = sum(f(r, "# Samples") for r in g["rows"])
    if ex < 0.004 * tot_exec and sm < 0.004 * tot_samp:
        continue
    wf = sum(f(r, "L1 Wavefronts Shared") for r in g["rows"])
    xs = sum(f(r, "L1 Wavefronts Shared Excessive") for r in g["rows"])
    ops = {}
    for r in g["rows"]:
        op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    st = {k: sum(f(r, k) for r in g["rows"]) for k in stall_keys}
    tops = " ".join(f"{k[6:]}:{100 * v / max(sm, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{g['start']:>6} {len(g['rows']):>5} {ex / len(g['rows']):>11.4g} {ex:>10.4g} {100 * ex / tot_exec:>6.1f} {100 * sm / max(tot_samp, 1):>6.1f} {wf:>10.4g} {xs:>10.4g}  {top} | {tops}")
