#!/usr/bin/env python
"""Log-mel kernel alone on a working set >> L2 (default 256 x 30 s): prints ms and achieved GB/s.  ncu target for the mel roofline.
    python tools/mel_prof.py [n_clips] [reps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_asr_b200 import B200AudioEncoder
from qwen3_asr_b200.synth import model_config, random_weights, speech_like

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = dict(model_config("1.7B")); cfg["encoder_layers"] = 1
enc = B200AudioEncoder(cfg, random_weights(cfg, seed=0), device=0, max_chunks=8)
n = 480000
base = torch.cat([torch.from_numpy(speech_like(n, i)) for i in range(8)]).cuda()
pcm = base.repeat(n_clips // 8)
offs = np.arange(n_clips + 1, dtype=np.int64) * n
for _ in range(2):
    enc.logmel_packed(pcm, offs)
torch.cuda.synchronize()
enc.profile(True)
for _ in range(reps):
    enc.logmel_packed(pcm, offs)
p = enc.profile_read()["logmel"]
ms = p["ms"] / p["launches"]
print(f"logmel {n_clips} x 30 s: {ms:.3f} ms/launch  {p['work'] / p['launches'] / ms / 1e6:.0f} GB/s  ({n_clips * 30 / ms / 1e3:.2f} M audio-s/s)")
enc.close()
