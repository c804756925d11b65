#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ (run once, in the build container).

The reference's arithmetic for this path lives in third-party classes that the
reference imports at run time: ``transformers``
``WhisperFeatureExtractor`` and ``Qwen3OmniMoeAudioEncoder`` (the class vLLM maps
Qwen3-ASR's audio_tower onto).  This script runs THOSE classes (transformers
5.5.0, torch CPU float32) and stores their outputs; the oracle restatement and
the CUDA path are then tested against the stored vectors without needing
/root/reference or a network at test time.

Real inputs: four of the reference's FLEURS fixtures
(/root/reference/E2Etest/data/audio/real/*.wav); only a 3 s excerpt of each is
stored (int16), next to the extractor's output for that excerpt.

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import os
import sys
import wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import CONFIGS, make_weights, noise_clip, speech_like  # noqa: E402
from oracle.encoder import token_len  # noqa: E402

REF_AUDIO = "/root/reference/E2Etest/data/audio/real"


def read_wav(path):
    with wave.open(path) as w:
        assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        return np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)


def hf_logmel(fe, x):
    out = fe(x, sampling_rate=16000, padding=True, truncation=False, return_attention_mask=True,
             return_tensors="np")
    feats = out["input_features"][0]
    n = int(out["attention_mask"][0].sum())
    return feats[:, :n].astype(np.float32)


def hf_encoder(cfg, weights, mels, attn_impl="eager"):
    """transformers module, float32 CPU, with the block-diagonal mask injected."""
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeAudioEncoder

    hc = Qwen3OmniMoeAudioEncoderConfig(
        num_mel_bins=cfg.num_mel_bins, encoder_layers=cfg.layers, encoder_attention_heads=cfg.heads,
        encoder_ffn_dim=cfg.ffn, d_model=cfg.d_model, output_dim=cfg.output_dim,
        n_window=cfg.n_window, n_window_infer=cfg.n_window_infer, conv_chunksize=500,
        downsample_hidden_size=cfg.downsample_hidden, max_source_positions=cfg.max_source_positions,
        activation_function="gelu", scale_embedding=False, dropout=0.0, attention_dropout=0.0,
        activation_dropout=0.0,
    )
    hc._attn_implementation = attn_impl
    enc = Qwen3OmniMoeAudioEncoder(hc).eval().float()
    missing, unexpected = enc.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    assert all("positional_embedding" in k for k in missing), missing

    # inject the window mask: layers accept attention_mask but forward never passes it
    for layer in enc.layers:
        orig = layer.forward

        def fwd(hidden_states, cu_seqlens, attention_mask=None, _orig=orig, **kw):
            mask = enc._prepare_attention_mask(hidden_states, cu_seqlens)
            return _orig(hidden_states, cu_seqlens, attention_mask=mask, **kw)

        layer.forward = fwd

    outs = []
    with torch.no_grad():
        for m in mels:  # one clip per call == the reference's one-request-per-job semantics
            feats = torch.from_numpy(np.ascontiguousarray(m))
            lens = torch.tensor([m.shape[1]], dtype=torch.long)
            outs.append(enc(feats, feature_lens=lens).last_hidden_state.float().numpy())
    return outs


def main():
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor(feature_size=128)
    gold = {}

    # ---- mel filter bank: every non-zero, as (bin, filter, value)
    fb = np.asarray(fe.mel_filters, dtype=np.float64)
    nz = np.nonzero(fb)
    gold["fb_rows"] = nz[0].astype(np.int16)
    gold["fb_cols"] = nz[1].astype(np.int16)
    gold["fb_vals"] = fb[nz]

    # ---- log-mel: synthetic + real excerpts
    clips = {
        "noise5s": noise_clip(80000, 0),
        "speech2s": speech_like(32000, 7),
        "odd": noise_clip(80077, 3),           # N % 160 != 0
        "short": speech_like(7200, 11),        # first WS trigger, 0.45 s
        "tone": (0.5 * np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000)).astype(np.float32),
        "zeros": np.zeros(4000, np.float32),
    }
    for name in ("english_01", "chinese_02", "thai_02", "hindi_01"):
        pcm = read_wav(os.path.join(REF_AUDIO, name + ".wav"))
        ex = pcm[16000 : 16000 + 48000]
        gold[f"pcm_{name}"] = ex
        clips[name] = ex.astype(np.float32) / 32768.0
    for name, x in clips.items():
        gold[f"mel_{name}"] = hf_logmel(fe, x)
    gold["mel_names"] = np.array(sorted(clips.keys()))

    # ---- token-length formula
    ts = np.arange(0, 3201)
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import _get_feat_extract_output_lengths
    gold["toklen_T"] = ts.astype(np.int32)
    gold["toklen"] = _get_feat_extract_output_lengths(torch.from_numpy(ts)).numpy().astype(np.int32)

    # ---- encoder: tiny dims on several lengths (windows, tails, short clip), and config 1 at 0.6B dims
    cfg = CONFIGS["tiny"]
    w = make_weights(cfg, seed=1)
    lens = [1056, 45, 300, 177, 100, 835]   # windows [104,33]; lone short clip; exact; tail; 1 chunk; 104+4
    mels = []
    for i, t in enumerate(lens):
        m = hf_logmel(fe, speech_like(t * 160, 50 + i))
        mels.append(m.astype(np.float32))
    outs = hf_encoder(cfg, w, mels)
    for i, (t, o) in enumerate(zip(lens, outs)):
        assert o.shape == (token_len(t), cfg.output_dim), (o.shape, t)
        gold[f"enc_tiny_{i}"] = o.astype(np.float32)
    gold["enc_tiny_lens"] = np.array(lens, np.int32)

    cfg = CONFIGS["0.6B"]
    w = make_weights(cfg, seed=2)
    m = hf_logmel(fe, noise_clip(80000, 0))
    m = torch.from_numpy(m).to(torch.bfloat16).float().numpy()   # what the tower sees in deployment
    o = hf_encoder(cfg, w, [m])[0]
    assert o.shape == (65, 1024)
    gold["enc_c1"] = o.astype(np.float32)

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **gold)
    sz = os.path.getsize(os.path.join(HERE, "golden.npz"))
    print(f"wrote golden.npz: {len(gold)} arrays, {sz/1e6:.2f} MB")


if __name__ == "__main__":
    main()
