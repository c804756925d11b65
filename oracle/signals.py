"""Synthetic 16 kHz inputs for the BASELINE.json configs (TEST INFRASTRUCTURE).

Definitions from SURVEY.md section 8(d); the "speech-like" formula follows the
reference's own generator, E2Etest/utils/audio.py:38-57, with a seeded
``default_rng`` instead of the global numpy RNG, and the WS pre-steps follow
src/server.py:26-29 (Butterworth band-pass) and :1335-1336 (int16 -> /32768).
"""

from __future__ import annotations

import numpy as np

SR = 16000


def noise_clip(n: int, seed: int = 0, amplitude: float = 0.1) -> np.ndarray:
    """Config 1: x = 0.1 * default_rng(seed).standard_normal(N), float32."""
    return (amplitude * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


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


def ws_prefilter(x: np.ndarray) -> np.ndarray:
    """int16 quantise -> /32768 -> 300-3400 Hz 4th-order Butterworth SOS (server.py:26-29,1335-1338)."""
    from scipy.signal import butter, sosfilt

    q = (np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16)
    f = q.astype(np.float32) / 32768.0
    sos = butter(4, [300, 3400], btype="bandpass", fs=SR, output="sos")
    return sosfilt(sos, f).astype(np.float32)


def config_clips(config: int, limit: int | None = None):
    """List of float32 clips for BASELINE.json configs[config-1] (1-based)."""
    clips = []
    if config == 1:
        clips = [noise_clip(80000, 0)]
    elif config == 2:
        n = 32 if limit is None else min(32, limit)
        clips = [speech_like(480000, i) for i in range(n)]
    elif config == 3:
        n = 128 if limit is None else min(128, limit)
        for i in range(n):
            ln = min(96000, 7200 * (1 + (5 * i) % 14))
            x = ws_prefilter(speech_like(ln, 1000 + i))
            if i % 4 == 3:
                x = np.concatenate([x, np.zeros(9600, np.float32)])
            if x.shape[0] < 8000:
                x = np.concatenate([x, np.zeros(8000 - x.shape[0], np.float32)])
            clips.append(x)
    elif config == 4:
        rng = np.random.default_rng(1234)
        total, j = 0, 0
        target = 3600 * SR
        while total < target and (limit is None or j < limit):
            ln = int(round(rng.uniform(1.0, 30.0) / 0.01)) * 160
            ln = min(ln, target - total)
            clips.append(speech_like(ln, 2000 + j))
            total += ln
            j += 1
    else:
        raise ValueError(f"unknown config {config}")
    return clips
