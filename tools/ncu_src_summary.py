#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: per kernel, an opcode histogram
(warp instructions executed) and the instructions with the most stall samples.
usage: ncu -i rep --page source --csv --print-source sass > src.csv; python tools/ncu_src_summary.py src.csv [kernel_idx] [top]"""
import csv, sys, collections

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
kernels = []
cur = None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
# ncu prints each kernel twice (two views); keep those whose Source column looks like SASS
for ki, k in enumerate(kernels):
    if which is not None and ki != which:
        continue
    h = k["hdr"]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    ops = collections.Counter(); samp = collections.Counter()
    tot_inst = 0; tot_samp = 0
    for r in k["rows"]:
        src = r[iS].strip()
        toks = src.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = ".".join(op.split(".")[:3])
        n = int(float(r[iE] or 0)); s = int(float(r[iN] or 0))
        ops[op] += n; samp[op] += s; tot_inst += n; tot_samp += s
    print(f"== [{ki}] {k['name'][:110]}\n   warp-instructions {tot_inst:,}  samples {tot_samp:,}")
    for op, n in ops.most_common(22):
        print(f"   {op:28s} {n:>14,} {100*n/max(tot_inst,1):5.1f}%   samples {100*samp[op]/max(tot_samp,1):5.1f}%")
    rows = sorted(k["rows"], key=lambda r: -int(float(r[iN] or 0)))[:top]
    print("   -- top stall lines")
    for r in rows:
        st = sorted(((int(float(r[i] or 0)), h[i]) for i in stall_cols), reverse=True)[:3]
        print(f"   {r[0][-6:]} {int(float(r[iN] or 0)):>7} {r[iS].strip()[:70]:70s} {' '.join(f'{c[6:]}={n}' for n, c in st if n)}")
    agg = collections.Counter()
    for r in k["rows"]:
        for i in stall_cols:
            agg[h[i][6:]] += int(float(r[i] or 0))
    print("   -- stall totals:", " ".join(f"{c}={n}" for c, n in agg.most_common(12)))
    print(f"   -- SASS lines: {len(k['rows'])} ({len(k['rows'])*16/1024:.0f} KB)")
