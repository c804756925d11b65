#!/usr/bin/env python
"""Which source lines execute a given SASS opcode (warp-instructions per frame), from an ncu report with source.
    python tools/ncu_op_lines.py report.ncu-rep frames OPCODE[,OPCODE...]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; frames = float(sys.argv[2]); want = sys.argv[3].split(',')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
h = None; cur = None; fname = None; agg = collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; continue
    if h is None or len(r) < 8: continue
    if r[0].isdigit():
        cur = f"{fname}:{r[0]} {r[1].strip()[:90]}"
        continue
    if r[0] == "" and r[2].startswith("0x"):
        tok = r[3].split()
        if not tok: continue
        op = tok[1] if tok[0].startswith('@') and len(tok) > 1 else tok[0]
        try: n = int(r[7])
        except ValueError: continue
        if op.split('.')[0] in want and n: agg[(cur, op)] += n
for (c, op), n in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
    print(f"{n / frames:7.2f} {op:16s} {c}")
