#!/usr/bin/env python
"""Determinism soak: the same batches encoded many times back to back (no host sync in between, programmatic dependent launch and
the pipelined host path active) must produce bit-identical outputs every time; different batch shapes are interleaved so that a
stale read of a reused workspace buffer would show.  Prints one JSON line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from qwen3_asr_b200 import B200AudioEncoder  # noqa: E402
from qwen3_asr_b200.synth import model_config, random_weights, speech_like  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    cfg = model_config("1.7B")
    enc = B200AudioEncoder(cfg, random_weights(cfg, seed=0))
    rng = np.random.default_rng(3)
    batches = []
    for b, n_clips in enumerate((32, 5, 17)):
        clips = [speech_like(int(rng.integers(8000, 480000)) if b else 480000, 10 * b + i) for i in range(n_clips)]
        batches.append(enc.pack_clips(clips))
    torch.cuda.synchronize()

    def checksum(t):
        return t.view(torch.int16).to(torch.int64).sum()

    ref = []
    for pcm, offs in batches:
        out, _ = enc.encode_pcm_packed(pcm, offs)
        ref.append(checksum(out).clone())
    sums = [[] for _ in batches]
    for r in range(reps):
        for k, (pcm, offs) in enumerate(batches):
            out, _ = enc.encode_pcm_packed(pcm, offs)
            sums[k].append(checksum(out))
    torch.cuda.synchronize()
    bad = [int(sum(int(s.item() != ref[k].item()) for s in sums[k])) for k in range(len(batches))]
    print(json.dumps({"reps": reps, "batches": [int(len(o) - 1) for _, o in batches], "mismatching_runs": bad, "ok": not any(bad)}))
    enc.close()
    if any(bad):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
