"""Synthetic weights and audio for benchmarks and smoke runs (no checkpoint / dataset is available offline).

Model dims: SURVEY.md appendix A.1.  Signals: the reference's own "speech-like" generator
(/root/reference/E2Etest/utils/audio.py:38-57) with a seeded RNG, as BASELINE.json's configs name them.
"""

from __future__ import annotations

import math

import numpy as np
import torch

SR = 16000

MODEL_DIMS = {
    "1.7B": dict(d_model=1024, encoder_layers=24, encoder_attention_heads=16, encoder_ffn_dim=4096, output_dim=2048),
    "0.6B": dict(d_model=896, encoder_layers=18, encoder_attention_heads=14, encoder_ffn_dim=3584, output_dim=1024),
    "tiny": dict(d_model=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=256, output_dim=192),
}
COMMON = dict(n_window=50, n_window_infer=800, downsample_hidden_size=480, num_mel_bins=128, max_source_positions=1500)


def model_config(name: str) -> dict:
    return {**MODEL_DIMS[name], **COMMON, "name": name}


def weight_shapes(cfg: dict) -> dict:
    c, d, f = cfg["downsample_hidden_size"], cfg["d_model"], cfg["encoder_ffn_dim"]
    shapes = {
        "conv2d1.weight": (c, 1, 3, 3), "conv2d1.bias": (c,),
        "conv2d2.weight": (c, c, 3, 3), "conv2d2.bias": (c,),
        "conv2d3.weight": (c, c, 3, 3), "conv2d3.bias": (c,),
        "conv_out.weight": (d, c * 16),
    }
    for i in range(cfg["encoder_layers"]):
        p = f"layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            shapes[p + f"self_attn.{nm}.weight"] = (d, d)
            shapes[p + f"self_attn.{nm}.bias"] = (d,)
        shapes[p + "self_attn_layer_norm.weight"] = (d,)
        shapes[p + "self_attn_layer_norm.bias"] = (d,)
        shapes[p + "fc1.weight"] = (f, d)
        shapes[p + "fc1.bias"] = (f,)
        shapes[p + "fc2.weight"] = (d, f)
        shapes[p + "fc2.bias"] = (d,)
        shapes[p + "final_layer_norm.weight"] = (d,)
        shapes[p + "final_layer_norm.bias"] = (d,)
    shapes.update({"ln_post.weight": (d,), "ln_post.bias": (d,), "proj1.weight": (d, d), "proj1.bias": (d,),
                   "proj2.weight": (cfg["output_dim"], d), "proj2.bias": (cfg["output_dim"],)})
    return shapes


def random_weights(cfg: dict, seed: int = 0) -> dict:
    """Seeded random-init weights (fan-in scaled so activations stay O(1)), bf16-representable float32."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in weight_shapes(cfg).items():
        if name.endswith("layer_norm.weight") or name == "ln_post.weight":
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        else:
            w = torch.randn(shape, generator=g) / math.sqrt(int(np.prod(shape[1:])))
            if name == "conv2d1.weight":
                w = w * 2.0
            if ".q_proj.weight" in name or ".k_proj.weight" in name:
                w = w * 3.0
        out[name] = w.to(torch.bfloat16).to(torch.float32).contiguous()
    return out


def speech_like(n: int, seed: int, peak: float | None = 0.9) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / SR
    sig = 0.05 * rng.standard_normal(n)
    for f in (150, 300, 600, 1200):
        sig += 0.1 * np.sin(2 * np.pi * f * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5 * t))
    sig *= 0.3 + 0.7 * np.abs(np.sin(2 * np.pi * 4 * t))
    if peak is not None:
        sig *= peak / max(np.abs(sig).max(), 1e-12)
    return sig.astype(np.float32)


def workload_c2(n_clips: int = 32, seconds: float = 30.0, seed0: int = 0):
    """BASELINE.json configs[1]: batch 32 x 30 s synthetic clips."""
    n = int(round(seconds * SR))
    return [speech_like(n, seed0 + i) for i in range(n_clips)]


def lpt_assign(frames, n_ranks: int):
    """Longest-processing-time-first assignment of clips to ranks by mel frames.
    Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(frames)), key=lambda i: (-int(frames[i]), i))
    loads = [0] * n_ranks
    bins = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += int(frames[i])
    for b in bins:
        b.sort()
    return bins


# ---- BASELINE.json configs as product-side synthetic workloads (bench.py --config c1 .. c5; definitions: SURVEY.md section 8(d)) ----
def noise_clip(n: int, seed: int = 0, amplitude: float = 0.1) -> np.ndarray:
    """C1: x = 0.1 * default_rng(seed).standard_normal(N), float32."""
    return (amplitude * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def workload_c1(seed: int = 0):
    """BASELINE.json configs[0]: one 5 s clip, batch 1 (0.6B dims)."""
    return [noise_clip(80000, seed)]


def workload_c3(seed0: int = 1000, n_streams: int = 128):
    """BASELINE.json configs[2]: one sliding window per WebSocket stream, as the socket delivers it: int16 PCM, length of stream i =
    min(96 000, 7 200 * (1 + (5 i mod 14))) samples (0.45 s .. 6 s), every 4th stream is a flush (gets 600 ms of silence appended and,
    in the dual-model variant, runs on the full model; src/server.py:1327-1329, 1351).  Returns (list of int16 arrays, flush flags)."""
    wins, flush = [], []
    for i in range(n_streams):
        n = min(96000, 7200 * (1 + (5 * i) % 14))
        x = speech_like(n, seed0 + i)
        wins.append((np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16))
        flush.append(i % 4 == 3)
    return wins, flush


def workload_c4_lengths(total_seconds: int = 3600, seed: int = 1234):
    """BASELINE.json configs[3]: segment lengths (samples) of one hour of audio cut into 1-30 s pieces (~230 segments)."""
    rng = np.random.default_rng(seed)
    lens, total, target = [], 0, total_seconds * SR
    while total < target:
        ln = min(int(round(rng.uniform(1.0, 30.0) / 0.01)) * 160, target - total)
        lens.append(ln)
        total += ln
    return lens


def workload_c4_clips(lens, indices, seed0: int = 2000):
    """The segments `indices` of the hour.  Segment j is a rotation of one 30 s speech-like signal (generating 3600 s of fresh
    signal would dominate the bench's start-up; the content does not change the timing)."""
    base = speech_like(30 * SR, seed0)
    return [np.roll(base, 997 * j)[: lens[j]].copy() for j in indices]
