#!/usr/bin/env python
"""Compact view of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f} ms)  launches {d['gpu_launches']}")
r = d["roofline"]; print(f"gemm family: {r['achieved']:.0f} TF/s frac {r['frac']:.3f} share {r['share_of_step']:.3f}")
m = d.get("roofline_mel")
if m: print(f"mel: {m['achieved']:.0f} GB/s frac {m['frac']:.3f}  {m['audio_s_per_s']/1e6:.2f} M audio-s/s  launch {m['avg_launch_ms']:.3f} ms")
print("clocks", d.get("clocks")); print("cpu", d.get("cpu_baseline"))
tot = 0
for k, v in d["kernels"].items():
    tot += v["ms_per_step"]
    tf = f"{v['tflops']:.0f} TF" if v.get("tflops") else ""
    print(f"  {k:18s} {v['ms_per_step']:7.3f} ms  x{v['launches_per_step']:.0f}  {tf}")
print(f"  sum {tot:.3f} ms (profiled step {d['ms_per_step_profiled']:.3f})")
