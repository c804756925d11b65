#!/usr/bin/env python
"""Summarise an ncu report of the log-mel kernel: headline metrics, stall mix, hottest source lines.
    python tools/ncu_mel_summary.py report.ncu-rep [frames]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 768000.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "sm__ctas_launched.sum")
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")
st = [(float(v), h.split("issue_stalled_")[1].split("_per_")[0]) for h, v in zip(hdr, vals) if "issue_stalled" in h and "per_issue_active" in h and v]
print("stalls per issue:", ", ".join(f"{n}:{x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
out, fname, h = [], None, None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        continue
    if r[0] == "Function Name" or h is None or r[0] == "":
        continue
    try:
        n, s = int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")])
    except ValueError:
        continue
    stall = {h[i][6:]: int(r[i] or 0) for i in range(len(h)) if h[i].startswith("stall_") and "Not Issued" not in h[i] and i < len(r) and (r[i] or "0").isdigit()}
    if n or s:
        out.append((n, s, fname, r[0], r[1].strip()[:90], stall))
tot, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"{tot / frames:.1f} warp-instr per frame; {ts} samples")
for n, s, f, l, text, stall in sorted(out, key=lambda x: -x[1])[:22]:
    top = ",".join(f"{a}:{b}" for a, b in sorted(stall.items(), key=lambda kv: -kv[1])[:3] if b)
    print(f"{n / frames:7.1f}/frame smp {100 * s / ts:5.1f}% {f}:{l} {text} | {top}")
