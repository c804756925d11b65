#!/usr/bin/env python
"""BASELINE.json configs[3]: one hour of synthetic audio, silence-split into ~230 segments of 1-30 s,
encoded by ONE process over all visible GPUs (B200EncoderPool: contiguous clip ranges of near-equal work, no collective).
Host buffers in and out.  One JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from qwen3_asr_b200 import B200EncoderPool  # noqa: E402
from qwen3_asr_b200.synth import model_config, random_weights, speech_like  # noqa: E402

SR = 16000


def main():
    n_dev = torch.cuda.device_count()
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cfg = model_config("1.7B")
    pool = B200EncoderPool(cfg, random_weights(cfg, seed=0), devices=list(range(n_dev)))
    rng = np.random.default_rng(1234)
    lens, total, target = [], 0, 3600 * SR
    while total < target:
        ln = min(int(round(rng.uniform(1.0, 30.0) / 0.01)) * 160, target - total)
        lens.append(ln)
        total += ln
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = torch.empty(int(offs[-1]), dtype=torch.float32, pin_memory=True)
    base = speech_like(30 * SR, 2000)            # the hour is tiled from one 30 s signal: content does not affect the timing
    for i, ln in enumerate(lens):
        pcm[int(offs[i]):int(offs[i + 1])] = torch.from_numpy(np.roll(base, 997 * i)[:ln])
    n_tok = int(sum(pool.lib.qasr_token_len(ln // 160) for ln in lens))
    out = torch.empty((n_tok, pool.output_dim), dtype=torch.bfloat16, pin_memory=True)

    def once():
        t0 = time.perf_counter()
        t, toks, devs = pool.submit_pcm_host(pcm, offs, out)
        pool.collect(t)
        return time.perf_counter() - t0, devs

    once()
    dts = [once()[0] for _ in range(reps)]
    _, devs = once()
    frames = np.array([ln // 160 for ln in lens])
    load = [int(frames[devs == d].sum()) for d in range(n_dev)]
    best = min(dts)
    print(json.dumps({"workload": "C4: 3600 s of audio in %d segments (1-30 s), 1.7B, one process, %d GPU(s), host buffers" % (len(lens), n_dev),
                      "n_gpus": n_dev, "segments": len(lens), "tokens": n_tok, "best_ms": best * 1e3, "median_ms": float(np.median(dts)) * 1e3,
                      "audio_s_per_s": 3600.0 / best, "frames_per_gpu": load,
                      "imbalance": max(load) / (sum(load) / n_dev) - 1.0}))
    pool.close()


if __name__ == "__main__":
    main()
