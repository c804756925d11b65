#!/usr/bin/env python
"""Executed warp-instructions per frame of the log-mel kernel, by SASS opcode and by source line, from an ncu report.
    python tools/ncu_mel_regions.py report.ncu-rep [frames]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 768000.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
ops = collections.Counter()
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "Address" or (h is None and "Source" in r):
        h = r
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        n = int(r[h.index("Instructions Executed")])
    except (ValueError, IndexError):
        continue
    s = r[h.index("Source")].strip()
    tok = s.split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    ops[op.split(".")[0]] += n
    tot += n
print(f"{tot / frames:.1f} warp-instr per frame")
for op, n in ops.most_common(40):
    print(f"{n / frames:8.1f}  {op}")
