#!/usr/bin/env python
"""BASELINE.json configs[0] on the GPU: one 5 s clip, batch 1 (0.6B and 1.7B) -- the latency-bound end of the path.
Prints one JSON line: wall time per call (device-resident PCM, sync included), the sum of the kernels' own durations, and the
weight-bandwidth floor (all encoder weights read once from HBM).  Kept under profiles/; not the headline bench."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from qwen3_asr_b200 import B200AudioEncoder  # noqa: E402
from qwen3_asr_b200.synth import model_config, random_weights  # noqa: E402

HBM_GBS = 6544.3


def main():
    out = {}
    x = (0.1 * np.random.default_rng(0).standard_normal(80000)).astype(np.float32)
    for name in ("0.6B", "1.7B"):
        cfg = model_config(name)
        w = random_weights(cfg, seed=0)
        weight_bytes = sum(int(np.prod(v.shape)) for k, v in w.items()) * 2
        enc = B200AudioEncoder(cfg, w, max_chunks=64)
        pcm, offs = enc.pack_clips([x])
        torch.cuda.synchronize()
        for _ in range(5):
            enc.encode_pcm_packed(pcm, offs)
        torch.cuda.synchronize()
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            enc.encode_pcm_packed(pcm, offs)
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            enc.encode_pcm_packed(pcm, offs)
        torch.cuda.synchronize()
        wall_async = (time.perf_counter() - t0) / reps
        enc.profile(True)
        enc.encode_pcm_packed(pcm, offs)
        prof = enc.profile_read()
        enc.profile(False)
        kern = sum(v["ms"] for v in prof.values())
        out[name] = {"wall_ms_per_call_synced": wall * 1e3, "wall_ms_per_call_back_to_back": wall_async * 1e3,
                     "kernel_ms_sum": kern, "launches": int(sum(v["launches"] for v in prof.values())),
                     "weight_bytes": weight_bytes, "weight_floor_ms": weight_bytes / (HBM_GBS * 1e9) * 1e3,
                     "audio_s_per_s_synced": 5.0 / wall}
        enc.close()
    print(json.dumps({"workload": "C1 on the GPU: one 5 s clip (80 000 samples), batch 1, PCM resident on the device", **out}))


if __name__ == "__main__":
    main()
