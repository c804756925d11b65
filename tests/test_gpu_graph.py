"""CUDA-graph replay of repeated call shapes (csrc/qasr.cu "graph cache"): bit-identical to the eager launch sequence, independent
of the caller's pointers, one graph per shape, bounded memory."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(monkeypatch, graph, cfg_name="tiny", **kw):
    from oracle import CONFIGS, make_weights
    from qwen3_asr_b200 import B200AudioEncoder

    monkeypatch.delenv("QASR_DEBUG_KEEP", raising=False)    # a handle that keeps debug copies runs eagerly
    if graph is None:
        monkeypatch.delenv("QASR_GRAPH", raising=False)
    else:
        monkeypatch.setenv("QASR_GRAPH", graph)
    cfg = CONFIGS[cfg_name]
    enc = B200AudioEncoder(cfg, make_weights(cfg, seed=1), **kw)
    monkeypatch.delenv("QASR_GRAPH", raising=False)
    return enc


def _clips(seed):
    from oracle.signals import speech_like

    return [speech_like(int(5.0 * 16000), seed), speech_like(int(2.45 * 16000), seed + 1), speech_like(int(11.3 * 16000) + 77, seed + 2)]


def test_graph_replay_equals_the_eager_path_bitwise(monkeypatch):
    eager = _make(monkeypatch, "0", max_chunks=64)
    graphed = _make(monkeypatch, None, max_chunks=64)
    try:
        assert eager.graph_stats()["graphs"] == 0
        for rep in range(4):                      # sighting 1: eager; 2: capture + first replay; 3, 4: replays -- with fresh tensors each time
            clips = _clips(10 * rep)              # same lengths, different audio: one shape
            want, t0 = eager.encode_pcm(clips)
            l0 = graphed.launch_count
            got, t1 = graphed.encode_pcm(clips)
            torch.cuda.synchronize()
            assert t0.tolist() == t1.tolist() and torch.equal(got, want), f"repetition {rep}"
            assert graphed.launch_count - l0 == eager.launch_count // (rep + 1), "a replay accounts for the kernels it launches"
        st = graphed.graph_stats()
        assert st["graphs"] == 1 and st["replays"] == 3 and st["bytes"] > 0, (st, graphed.lib.qasr_last_error())
        assert eager.graph_stats() == {"graphs": 0, "replays": 0, "bytes": 0}
        # the mel entry point (what the server hook calls), float32 and bf16 features, strided input
        for dtype in (torch.float32, torch.bfloat16):
            for rep in range(3):
                mel, flens = eager.logmel(_clips(100 + rep))
                wide = torch.zeros((128, mel.shape[1] + 24), dtype=dtype, device=mel.device)
                wide[:, : mel.shape[1]] = mel.to(dtype)
                want = eager.encode(wide, flens)
                got = graphed.encode(wide, flens)
                torch.cuda.synchronize()
                assert torch.equal(got, want), (dtype, rep)
        assert graphed.graph_stats()["graphs"] == 3
        # a different shape is its own graph; an empty clip in the batch is part of the shape
        for rep in range(3):
            clips = _clips(200 + rep)[:2] + [np.zeros(0, np.float32)]
            want, _ = eager.encode_pcm(clips)
            got, toks = graphed.encode_pcm(clips)
            torch.cuda.synchronize()
            assert torch.equal(got, want) and toks[2] == 0
        assert graphed.graph_stats()["graphs"] == 4
    finally:
        eager.close()
        graphed.close()


def test_graph_replay_on_other_streams_and_after_other_shapes(monkeypatch):
    """A graph is replayed on whatever stream the caller is on, interleaved with eager calls of other shapes on the same workspaces."""
    from oracle.signals import speech_like

    eager = _make(monkeypatch, "0", max_chunks=64)
    graphed = _make(monkeypatch, None, max_chunks=64)
    try:
        a = [speech_like(3 * 16000, 1)]
        b = [speech_like(7 * 16000, 2), speech_like(16000, 3)]
        want_a, want_b = eager.encode_pcm(a)[0].clone(), eager.encode_pcm(b)[0].clone()
        streams = [torch.cuda.Stream() for _ in range(2)]
        outs = []
        for rep in range(6):
            with torch.cuda.stream(streams[rep % 2]):
                outs.append(("a", graphed.encode_pcm(a)[0]))
                outs.append(("b", graphed.encode_pcm(b)[0]))
        torch.cuda.synchronize()
        for name, o in outs:
            assert torch.equal(o, want_a if name == "a" else want_b)
        assert graphed.graph_stats()["graphs"] == 2 and graphed.graph_stats()["replays"] == 10
    finally:
        eager.close()
        graphed.close()


def test_large_calls_stay_eager_unless_asked(monkeypatch):
    from oracle.signals import speech_like

    enc = _make(monkeypatch, None, max_chunks=256)
    every = _make(monkeypatch, "all", max_chunks=256)
    try:
        clips = [speech_like(30 * 16000, i) for i in range(5)]     # 150 one-second chunks > 128
        outs = [enc.encode_pcm(clips)[0].clone() for _ in range(3)]
        outs2 = [every.encode_pcm(clips)[0].clone() for _ in range(3)]
        torch.cuda.synchronize()
        assert enc.graph_stats()["graphs"] == 0 and every.graph_stats() ["graphs"] == 1 and every.graph_stats()["replays"] == 2
        for o in outs + outs2:
            assert torch.equal(o, outs[0])
    finally:
        enc.close()
        every.close()


def test_real_width_single_window_graph(monkeypatch):
    """BASELINE config 1 shape (one 5 s clip) at 0.6B dims: the graph path equals the eager path bitwise."""
    from oracle.signals import noise_clip

    eager = _make(monkeypatch, "0", "0.6B", max_chunks=64)
    graphed = _make(monkeypatch, None, "0.6B", max_chunks=64)
    try:
        clip = [noise_clip(80000, 0)]
        want = eager.encode_pcm(clip)[0]
        for _ in range(3):
            got = graphed.encode_pcm(clip)[0]
        torch.cuda.synchronize()
        assert torch.equal(got, want) and graphed.graph_stats()["replays"] == 2
    finally:
        eager.close()
        graphed.close()


@pytest.mark.parametrize("model", ["0.6B", "1.7B"])
def test_small_m_path_is_bit_identical(monkeypatch, model):
    """Calls of <= 128 tokens (one window / one chunk: what the unbatched reference sends, src/server.py:79-94) take the small-M GEMM
    kernel (tc_gemm_small.cuh: one CTA per 16/32-column weight slice, the whole slice requested up front).  Same K order, same
    epilogue code: bit-identical to the persistent pair kernel (QASR_SMALL_M=0), so a window alone still equals the window in a
    batch -- checked at the real model widths for 1 .. 9 one-second chunks, eager and graph-replayed."""
    from oracle.signals import speech_like
    from qwen3_asr_b200 import B200AudioEncoder
    from qwen3_asr_b200.synth import model_config, random_weights

    cfg = dict(model_config(model))
    cfg["encoder_layers"] = 3
    w = random_weights(cfg, seed=3)
    monkeypatch.setenv("QASR_SMALL_M", "0")
    big = B200AudioEncoder(cfg, w, max_chunks=64)
    monkeypatch.delenv("QASR_SMALL_M")
    small = B200AudioEncoder(cfg, w, max_chunks=64)
    try:
        for i, secs in enumerate([0.45, 1.0, 2.45, 5.0, 6.0, 8.99, 9.84]):      # 6 .. 128 tokens
            clip = [speech_like(int(secs * 16000), 40 + i)]
            want, t0 = big.encode_pcm(clip)
            for rep in range(3):                                                 # eager, capture, replay
                got, t1 = small.encode_pcm(clip)
                torch.cuda.synchronize()
                assert t0.tolist() == t1.tolist() and int(t0[0]) <= 128
                assert torch.equal(got, want), (model, secs, rep)
        # and inside a batch (pair kernel on both handles) the same clip gives the same rows
        clips = [speech_like(int(s * 16000), 40 + i) for i, s in enumerate([0.45, 1.0, 2.45, 5.0, 6.0, 8.99, 9.84])]
        both, toks = small.encode_pcm(clips)
        alone, _ = small.encode_pcm([clips[3]])
        torch.cuda.synchronize()
        o = int(sum(toks[:3]))
        assert torch.equal(both[o:o + int(toks[3])], alone)
    finally:
        big.close()
        small.close()
