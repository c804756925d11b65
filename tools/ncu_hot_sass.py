#!/usr/bin/env python
"""Hottest SASS instructions (by stall samples) of an ncu report with source, with the CUDA line each belongs to and its neighbours.
    python tools/ncu_hot_sass.py report.ncu-rep [top_n] [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src)) if r]
h = next(r for r in rows if r[0] == "Address")
iS, iN, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
st = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ins = []
for r in rows:
    if r[0].startswith("0x") and len(r) >= len(h) - 2:
        try: ins.append((int(r[iS]), int(r[iN]), r[iSrc].strip(), {h[i][6:]: int(r[i] or 0) for i in st}))
        except ValueError: pass
tot = sum(i[0] for i in ins)
order = sorted(range(len(ins)), key=lambda i: -ins[i][0])[:top]
for i in order:
    s, n, text, stall = ins[i]
    topst = ",".join(f"{a}:{b}" for a, b in sorted(stall.items(), key=lambda kv: -kv[1])[:2] if b)
    print(f"--- {100*s/tot:.1f}% [{topst}]")
    for k in range(max(0, i - ctx), min(len(ins), i + 2)):
        print(f"   {'>>' if k == i else '  '} {ins[k][0]:6d} {ins[k][2][:110]}")
