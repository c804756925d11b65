#!/usr/bin/env python
"""Per-CUDA-source-line warp-stall summary from an ncu report.

    python tools/ncu_lines.py report.ncu-rep [top_n]

Runs `ncu -i report --page source --csv --print-source cuda,sass` and aggregates samples per source line."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel, fname, hdr = None, None, None
    per_kernel = {}
    for r in rows:
        if not r:
            continue
        if r[0] in ("Kernel Name", "Function Name"):
            kernel = r[1]
            continue
        if r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or len(r) < len(hdr) - 2:
            continue  # SASS rows / plain source listing
        i_s = hdr.index("# Samples")
        def num(v):
            try:
                return int(v)
            except ValueError:
                return 0

        stall = {hdr[i][6:]: num(r[i]) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i] and i < len(r)}
        n = num(r[i_s])
        if n == 0:
            continue
        per_kernel.setdefault(kernel, []).append((n, fname, r[0], r[1].strip()[:95], stall))
    for k, lines in per_kernel.items():
        tot = sum(l[0] for l in lines)
        print(f"===== {k[:150]}\n      total samples {tot}")
        for n, f, ln, src, stall in sorted(lines, key=lambda l: -l[0])[:top_n]:
            top = ", ".join(f"{a}:{b}" for a, b in sorted(stall.items(), key=lambda kv: -kv[1])[:3] if b)
            print(f"{100*n/tot:5.1f}% {f}:{ln:>4} {src:95s} | {top}")


if __name__ == "__main__":
    main()
