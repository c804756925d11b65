#!/usr/bin/env python
"""Are two builds of the library bit-identical on the same input?  (child processes: one library per process)
    python tools/ab_bits.py libA.so libB.so [model]"""
import os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from qwen3_asr_b200 import B200AudioEncoder
from qwen3_asr_b200.synth import model_config, random_weights, speech_like
cfg = model_config(sys.argv[1])
enc = B200AudioEncoder(cfg, random_weights(cfg, seed=0), device=0, max_chunks=64)
clips = [speech_like(n, 5 + i) for i, n in enumerate([16000 * 5, 16000 * 30, 7200, 16000 * 11 + 77])]
out, toks = enc.encode_pcm(clips)
torch.cuda.synchronize()
print(hashlib.sha256(out.cpu().view(torch.int16).numpy().tobytes()).hexdigest(), [int(t) for t in toks])
''' % ROOT
model = sys.argv[3] if len(sys.argv) > 3 else "0.6B"
res = []
for lib in sys.argv[1:3]:
    env = dict(os.environ, QASR_B200_LIB=os.path.abspath(lib), QASR_GRAPH="0")
    r = subprocess.run([sys.executable, "-c", CHILD, model], env=env, capture_output=True, text=True)
    print(lib, r.stdout.strip() or r.stderr[-400:])
    res.append(r.stdout.strip())
print("IDENTICAL" if res[0] and res[0] == res[1] else "DIFFERENT")
