#!/usr/bin/env python
"""BASELINE.json configs[2] -- "WebSocket sliding-window re-encode: 128 concurrent streams x 6 s context, 1.7B, varlen batch".

Measures, on one GPU, PCM16 bytes (host) -> [int16 -> /32768 -> 300-3400 Hz band-pass -> flush pad] -> log-mel -> encoder for the
128 windows of SURVEY section 8(d) C3, (a) as ONE ragged batch through B200PreFrontend.encode_windows (what WindowBatcher
forms) and (b) one window per call (what the reference does: one job per window, src/server.py:79-94).  Prints one JSON line.
Not the headline bench (that is bench.py on C2); a parity-config measurement kept under profiles/."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from qwen3_asr_b200 import B200AudioEncoder, B200PreFrontend  # noqa: E402
from qwen3_asr_b200.synth import model_config, random_weights, speech_like  # noqa: E402


def main():
    cfg = model_config("1.7B")
    enc = B200AudioEncoder(cfg, random_weights(cfg, seed=0), max_chunks=1024)
    pre = B200PreFrontend(enc)
    wins, flush = [], []
    for i in range(128):
        n = min(96000, 7200 * (1 + (5 * i) % 14))
        x = speech_like(n, 1000 + i)
        wins.append((np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16).tobytes())
        flush.append(i % 4 == 3)
    audio_s = sum(len(w) / 2 + (9600 if f else 0) for w, f in zip(wins, flush)) / 16000.0

    def batched():
        hid, toks = pre.encode_windows(wins, 16000, pad_silence=flush)
        torch.cuda.synchronize()
        return hid, toks

    def per_window():
        outs = []
        for w, f in zip(wins, flush):
            outs.append(pre.encode_windows([w], 16000, pad_silence=[f])[0])
        torch.cuda.synchronize()
        return outs

    hid, toks = batched()
    singles = per_window()
    same = torch.equal(hid, torch.cat(singles, 0))          # batch invariance, end to end through the pre-frontend
    res = {}
    for name, fn, reps in (("batched", batched, 20), ("per_window", per_window, 3)):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        res[name] = (time.perf_counter() - t0) / reps
    enc.profile(True)
    batched()
    prof = enc.profile_read()
    enc.profile(False)
    print(json.dumps({
        "workload": "C3: 128 WS windows (0.45-6 s, every 4th +600 ms flush silence), PCM16 bytes on the host -> tokens on the device, 1.7B bf16",
        "audio_s": audio_s, "tokens": int(toks.sum()),
        "batched_ms": res["batched"] * 1e3, "batched_audio_s_per_s": audio_s / res["batched"],
        "per_window_ms_total": res["per_window"] * 1e3, "per_window_audio_s_per_s": audio_s / res["per_window"],
        "per_window_ms_each": res["per_window"] * 1e3 / 128,
        "batched_equals_per_window_bitwise": bool(same),
        "kernels_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
    }))


if __name__ == "__main__":
    main()
