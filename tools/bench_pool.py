#!/usr/bin/env python
"""One process, all GPUs of the box: B200EncoderPool (qasr_pool_*) on the C2 workload scaled to the pool
(32 x 30 s clips per GPU per batch), host buffers in and out, batches pipelined two deep.  One JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from qwen3_asr_b200 import B200EncoderPool  # noqa: E402
from qwen3_asr_b200.synth import model_config, random_weights, speech_like  # noqa: E402


def main():
    n_dev = torch.cuda.device_count()
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    cfg = model_config("1.7B")
    pool = B200EncoderPool(cfg, random_weights(cfg, seed=0), devices=list(range(n_dev)))
    base = [speech_like(480000, i) for i in range(32)]
    n_clips = 32 * n_dev
    offs = np.arange(n_clips + 1, dtype=np.int64) * 480000
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3     # batches in flight pool-wide (2 leave ~1 ms of D2H + collect/submit round trip exposed per step)
    bufs = []
    for b in range(max(2, depth)):
        pcm = torch.empty(n_clips * 480000, dtype=torch.float32, pin_memory=True)
        for i in range(n_clips):
            pcm[i * 480000:(i + 1) * 480000] = torch.from_numpy(base[(i + b) % 32])
        out = torch.empty((n_clips * 390, pool.output_dim), dtype=torch.bfloat16, pin_memory=True)
        bufs.append((pcm, out))

    def run(k, depth=2):
        tickets = []
        t0 = time.perf_counter()
        for s in range(k):
            pcm, out = bufs[s % len(bufs)]
            if len(tickets) == depth:
                pool.collect(tickets.pop(0))
            tickets.append(pool.submit_pcm_host(pcm, offs, out)[0])
        for t in tickets:
            pool.collect(t)
        return time.perf_counter() - t0

    run(3, depth)
    pool.stats(reset=True)
    dt = run(steps, depth)
    stats = pool.stats()
    audio_s = n_clips * 30.0 * steps
    _, toks, devs = pool.submit_pcm_host(bufs[0][0], offs, bufs[0][1])
    print(json.dumps({"workload": f"C2 x {n_dev} GPUs in ONE process (B200EncoderPool), host buffers, {depth} batches in flight",
                      "n_gpus": n_dev, "steps": steps, "ms_per_step": dt / steps * 1e3, "e2e_audio_s_per_s": audio_s / dt,
                      "clips_per_device": np.bincount(devs, minlength=n_dev).tolist(), "batches_in_flight": depth,
                      "per_member_ms": stats}))
    pool.close()


if __name__ == "__main__":
    main()
