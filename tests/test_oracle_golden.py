"""The oracle restatement against the golden vectors produced by the in-container
transformers classes (tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest
import torch

from conftest import range_rel
from oracle import CONFIGS, encoder_forward, logmel, logmel_torch_f32, make_weights, mel_filter_bank, token_len
from oracle.encoder import chunk_plan, window_lens
from oracle.signals import noise_clip, speech_like


def test_filter_bank_exact(golden):
    fb = mel_filter_bank()
    ref = np.zeros_like(fb)
    ref[golden["fb_rows"], golden["fb_cols"]] = golden["fb_vals"]
    assert fb.shape == (201, 128)
    assert np.count_nonzero(fb) == 394
    assert np.abs(fb - ref).max() == 0.0
    assert (np.count_nonzero(fb, axis=1) <= 2).all()


def _clip(golden, name):
    if f"pcm_{name}" in golden:
        return golden[f"pcm_{name}"].astype(np.float32) / 32768.0
    return {
        "noise5s": lambda: noise_clip(80000, 0),
        "speech2s": lambda: speech_like(32000, 7),
        "odd": lambda: noise_clip(80077, 3),
        "short": lambda: speech_like(7200, 11),
        "tone": lambda: (0.5 * np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000)).astype(np.float32),
        "zeros": lambda: np.zeros(4000, np.float32),
    }[name]()


def test_logmel_vs_golden(golden):
    for name in golden["mel_names"]:
        x = _clip(golden, str(name))
        ref = golden[f"mel_{name}"]
        got = logmel(x)
        assert got.shape == ref.shape == (128, x.shape[0] // 160), name
        # the golden (torch f32 STFT) carries ~2e-5 abs of its own rounding noise
        assert range_rel(got, ref) <= 1e-4, (name, range_rel(got, ref))
        got32 = logmel_torch_f32(x)
        assert np.abs(got32 - ref).max() <= 1e-6, name


def test_logmel_zero_clip(golden):
    assert np.all(logmel(np.zeros(4000, np.float32)) == -1.5)
    assert np.all(golden["mel_zeros"] == -1.5)


def test_token_len_formula(golden):
    for t, n in zip(golden["toklen_T"], golden["toklen"]):
        assert token_len(int(t)) == int(n)
    assert token_len(500) == 65 and token_len(600) == 78 and token_len(3000) == 390


def test_chunk_plan_and_windows():
    cfg = CONFIGS["1.7B"]
    assert chunk_plan(250) == [(0, 100, 100), (100, 100, 100), (200, 50, 100)]
    assert chunk_plan(45) == [(0, 45, 45)]
    assert window_lens(390, cfg, 3000) == [104, 104, 104, 78]
    assert window_lens(token_len(1056), cfg, 1056) == [104, 33]
    assert window_lens(token_len(45), cfg, 45) == [6]


def test_encoder_tiny_vs_golden(golden):
    cfg = CONFIGS["tiny"]
    w = make_weights(cfg, seed=1)
    lens = golden["enc_tiny_lens"]
    for i, t in enumerate(lens):
        m = logmel_torch_f32(speech_like(int(t) * 160, 50 + i))
        out, toks = encoder_forward(w, cfg, [m])
        ref = golden[f"enc_tiny_{i}"]
        assert toks == [ref.shape[0]]
        assert range_rel(out.numpy(), ref) <= 2e-5, (i, int(t), range_rel(out.numpy(), ref))


def test_encoder_batch_equals_per_clip():
    """Standalone per-clip semantics: batching must not change any clip's tokens."""
    cfg = CONFIGS["tiny"]
    w = make_weights(cfg, seed=1)
    mels = [logmel(speech_like(t * 160, 90 + i)) for i, t in enumerate((77, 300, 1056))]
    out, toks = encoder_forward(w, cfg, mels)
    s = 0
    for m, n in zip(mels, toks):
        alone, _ = encoder_forward(w, cfg, [m])
        assert torch.equal(out[s : s + n], alone) or (out[s : s + n] - alone).abs().max() < 1e-5
        s += n


@pytest.mark.slow
def test_encoder_config1_vs_golden(golden):
    cfg = CONFIGS["0.6B"]
    w = make_weights(cfg, seed=2)
    m = logmel_torch_f32(noise_clip(80000, 0))
    m = torch.from_numpy(m).to(torch.bfloat16).float().numpy()
    out, toks = encoder_forward(w, cfg, [m])
    assert toks == [65]
    assert range_rel(out.numpy(), golden["enc_c1"]) <= 2e-5
