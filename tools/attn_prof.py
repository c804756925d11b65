#!/usr/bin/env python
"""Window attention kernel alone at the C2 shape (128 windows of 104/78 tokens x 16 heads): ncu target.
    python tools/attn_prof.py [impl: tc|mma] [reps]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_asr_b200 import load_library
from qwen3_asr_b200._lib import check

impl = sys.argv[1] if len(sys.argv) > 1 else "tc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = load_library()
heads, d = 16, 1024
lens = [104, 104, 104, 78] * 32
n = sum(lens)
qkv = (torch.randn(n, 3 * d, device="cuda") * 1.5).to(torch.bfloat16)
out = torch.empty(n, d, dtype=torch.bfloat16, device="cuda")
wins, s = [], 0
for wl in lens:
    wins += [s, wl]
    s += wl
wh = (C.c_int32 * len(wins))(*wins)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
def run():
    if impl == "tc":
        check(lib, lib.qasr_debug_attention_tc(p(qkv), p(out), wh, len(lens), n, d, heads, st), "tc")
    else:
        check(lib, lib.qasr_debug_attention(p(qkv), p(out), wh, len(lens), d, heads, st), "mma")
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
print(f"{impl}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call (includes a small H2D + sync per call)")
