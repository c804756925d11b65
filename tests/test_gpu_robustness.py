"""Isolation and misuse: what must hold when one client's input is bad or the API is driven outside its happy path.
All through the C ABI on a real device (round-1 advisor findings)."""

import ctypes as C
import os
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    from oracle import CONFIGS, make_weights
    from qwen3_asr_b200 import B200AudioEncoder

    cfg = CONFIGS["tiny"]
    w = make_weights(cfg, seed=1)
    enc = B200AudioEncoder(cfg, w, max_chunks=64)
    yield cfg, w, enc
    enc.close()


def _clips():
    from oracle.signals import speech_like

    # lengths chosen so that windows end off a multiple of 16 tokens: 3.3 s -> 43 tokens, 11 s -> 104 + 39, 0.7 s -> 10
    return [speech_like(int(3.3 * 16000), 1), speech_like(11 * 16000, 2), speech_like(int(0.7 * 16000), 3), speech_like(5 * 16000, 4)]


def test_a_nan_clip_does_not_leak_into_its_batch_neighbours(tiny):
    """The tcgen05 attention multiplies P V over keys rounded up to 16: V rows past a window's end (the NEXT clip's rows, or stale
    rows of an earlier, larger batch) are zeroed on chip, so a NaN / Inf there cannot turn 0 * x into NaN.  Every clean clip must
    come out bit-identical to the batch without the bad clip, in this call and in the (smaller) calls that follow.  Poisoned at the
    feature level (non-finite log-mel columns reach the encoder as they are) and at the PCM level (the log-mel clamp is an fmax,
    which drops NaN: the bad clip may even come out finite -- the neighbours must not move either way)."""
    _, _, enc = tiny
    clips = _clips()
    mel, flens = enc.logmel(clips)
    ref = enc.encode(mel, flens).clone()
    toks = enc.last_token_lens.copy()
    torch.cuda.synchronize()
    assert torch.isfinite(ref.float()).all()
    offs = np.concatenate([[0], np.cumsum(toks)])
    cols = np.concatenate([[0], np.cumsum(flens)])
    for bad_i, poison in ((1, float("nan")), (0, float("inf")), (2, float("nan")), (3, float("-inf"))):
        bad = mel.clone()
        bad[:, int(cols[bad_i]) + 3:int(cols[bad_i]) + 9] = poison
        out = enc.encode(bad, flens)
        torch.cuda.synchronize()
        for i in range(len(clips)):
            got, want = out[offs[i]:offs[i + 1]], ref[offs[i]:offs[i + 1]]
            if i == bad_i:
                assert not torch.isfinite(got.float()).all()    # the poisoned clip itself is garbage, as it is in the reference
            else:
                assert torch.equal(got, want), f"clip {i} changed when clip {bad_i} was poisoned"
        # a later, smaller batch reuses the workspace rows the poisoned clip left behind
        small = enc.encode(mel[:, :int(cols[1])].contiguous(), flens[:1])
        torch.cuda.synchronize()
        assert torch.equal(small, ref[offs[0]:offs[1]])
    ref_pcm, _ = enc.encode_pcm(clips)
    for bad_i, poison in ((1, np.nan), (0, np.inf)):
        bad = [c.copy() for c in clips]
        bad[bad_i][1000:1100] = poison
        out, toks2 = enc.encode_pcm(bad)
        torch.cuda.synchronize()
        assert toks2.tolist() == toks.tolist()
        for i in range(len(clips)):
            if i != bad_i:
                assert torch.equal(out[offs[i]:offs[i + 1]], ref_pcm[offs[i]:offs[i + 1]]), f"clip {i} changed when clip {bad_i}'s PCM was poisoned"


def test_an_empty_window_does_not_fail_its_batch(tiny):
    """qasr_ws_window keeps an empty WS window empty (the reference returns '' for it, server.py:1331-1332); the encoder entry
    points accept the zero-length clip: 0 frames, 0 tokens, everyone else unaffected."""
    from qwen3_asr_b200 import B200PreFrontend, QasrError

    _, _, enc = tiny
    pre = B200PreFrontend(enc)
    rng = np.random.default_rng(5)
    wins = [(rng.standard_normal(n) * 3000).astype(np.int16) for n in (16000, 0, 9000)]
    hid, toks = pre.encode_windows(wins, 16000, pad_silence=[False, False, True])
    torch.cuda.synchronize()
    assert toks[1] == 0 and toks[0] > 0 and toks[2] > 0
    alone0, _ = pre.encode_windows([wins[0]], 16000, pad_silence=[False])
    alone2, _ = pre.encode_windows([wins[2]], 16000, pad_silence=[True])
    torch.cuda.synchronize()
    assert torch.equal(hid[: toks[0]], alone0) and torch.equal(hid[toks[0]:], alone2)
    # float entry points: an empty clip in the middle, and a batch of only empty clips
    clips = _clips()[:2]
    ref, t_ref = enc.encode_pcm(clips)
    out, t = enc.encode_pcm([clips[0], np.zeros(0, np.float32), clips[1]])
    torch.cuda.synchronize()
    assert t.tolist() == [t_ref[0], 0, t_ref[1]] and torch.equal(out, ref)
    out, t = enc.encode_pcm([np.zeros(0, np.float32)])
    assert out.shape[0] == 0 and t.tolist() == [0]
    mel, flens = enc.logmel([np.zeros(0, np.float32), clips[0]])
    assert flens.tolist() == [0, len(clips[0]) // 160]
    # 1..200 samples cannot be reflect-padded (torch.stft raises too): still a per-call error, with the clip named
    with pytest.raises(QasrError, match="clip 1 has 150"):
        enc.encode_pcm([clips[0], np.zeros(150, np.float32)])


def test_empty_windows_in_the_batcher_resolve_without_an_encode(tiny):
    from qwen3_asr_b200 import B200PreFrontend
    from qwen3_asr_b200.batcher import WindowBatcher

    _, _, enc = tiny
    pre = B200PreFrontend(enc)
    seen = []

    def encode(w, f):
        seen.append(len(w))
        return pre.encode_windows(w, 16000, pad_silence=f)

    b = WindowBatcher(encode, max_wait_ms=100.0)
    rng = np.random.default_rng(6)
    full = (rng.standard_normal(16000) * 3000).astype(np.int16).tobytes()
    futs = [b.submit(full), b.submit(b""), b.submit(full, flush=True)]
    outs = [f.result(timeout=30) for f in futs]
    b.close()
    assert seen == [2] and outs[1].shape[0] == 0 and outs[0].shape[0] > 0


def test_calls_on_different_streams_are_ordered(tiny):
    """One set of workspaces per handle: a call arriving on another stream than the previous one waits for it (the hook reads
    torch.cuda.current_stream(), which differs between the server's _cuda_stream and the default stream)."""
    _, _, enc = tiny
    clips = _clips()
    ref = [enc.encode_pcm([c])[0].clone() for c in clips]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for rep in range(6):
        for i, c in enumerate(clips):
            with torch.cuda.stream(streams[(rep + i) % 3]):
                outs.append((i, enc.encode_pcm([c])[0]))
    torch.cuda.synchronize()
    for i, o in outs:
        assert torch.equal(o, ref[i])


def test_more_than_two_unwaited_submits_is_an_error(tiny):
    from qwen3_asr_b200 import QasrError

    _, _, enc = tiny
    c = _clips()[0]
    pcm = torch.from_numpy(c).pin_memory()
    offs = np.array([0, len(c)], dtype=np.int64)
    outs = [torch.empty((64, enc.output_dim), dtype=torch.bfloat16).pin_memory() for _ in range(3)]
    t1, _ = enc.submit_pcm_host(pcm, offs, outs[0])
    t2, _ = enc.submit_pcm_host(pcm, offs, outs[1])
    with pytest.raises(QasrError, match="un-waited"):
        enc.submit_pcm_host(pcm, offs, outs[2])
    enc.wait(t1)
    t3, toks = enc.submit_pcm_host(pcm, offs, outs[2])
    enc.wait(t2)
    enc.wait(t3)
    n = int(toks[0])
    assert torch.equal(outs[0][:n], outs[1][:n]) and torch.equal(outs[0][:n], outs[2][:n])


def test_a_mistyped_switch_is_an_error_not_a_silent_default(monkeypatch):
    from oracle import CONFIGS, make_weights
    from qwen3_asr_b200 import B200AudioEncoder, QasrError

    cfg = CONFIGS["tiny"]
    w = make_weights(cfg, seed=1)
    for name, bad in (("QASR_ATTENTION", "mma-sync"), ("QASR_LN", "unfolded"), ("QASR_DEBUG_SIMT", "yes"), ("QASR_PDL", "off")):
        monkeypatch.setenv(name, bad)
        with pytest.raises(QasrError, match=name):
            B200AudioEncoder(cfg, w, max_chunks=16)
        monkeypatch.delenv(name)
    monkeypatch.setenv("QASR_ATTENTION", "mma_sync")
    B200AudioEncoder(cfg, w, max_chunks=16).close()


def test_queue_binding_over_the_real_encoder(tiny, monkeypatch):
    """queue_binding.QueueBinding with its default encode (log-mel + encoder on the collector thread's own stream): windows
    submitted together are encoded as one ragged batch and every job receives bit-exactly what its own call would have produced."""
    import asyncio
    import types

    from qwen3_asr_b200 import server_hook
    from qwen3_asr_b200.queue_binding import QueueBinding
    from test_queue_binding import FakeTower, StandInQueue

    _, _, enc = tiny
    monkeypatch.setenv("B200_ENCODER", "1")
    server_hook.unload()
    tower = FakeTower()

    class Model:
        model = types.SimpleNamespace(thinker=types.SimpleNamespace(audio_tower=tower))

        def transcribe(self, audio_sr, language=None, return_time_stamps=False):
            audio, _ = audio_sr
            mel, flens = enc.logmel([audio])
            out = self.model.thinker.audio_tower.forward(mel, feature_lens=flens)
            return [types.SimpleNamespace(text="", language="en", hidden=out.last_hidden_state)]

    m = Model()
    server = types.SimpleNamespace(model=m, _fast_model=None, _infer_queue=StandInQueue(), USE_SPECULATIVE=False)
    server._do_transcribe = lambda audio, sr, lang, ts, use_fast=False: m.transcribe((audio, sr))
    server_hook.install(server, batch_windows=False)
    server_hook.try_load_b200_encoder(m, factory=lambda t: enc)
    binding = QueueBinding(server, max_wait_ms=50.0)
    binding.install()
    clips = _clips() + [c[::-1].copy() for c in _clips()]

    async def one(audio):
        sr, lang_code = 16000, None
        res = await server._infer_queue.submit(lambda: server._do_transcribe(audio, sr, lang_code, False), priority=0)
        return res[0].hidden

    async def main():
        server._infer_queue.start()
        return await asyncio.gather(*[one(c) for c in clips])

    outs = asyncio.run(main())
    torch.cuda.synchronize()
    batches = list(binding._batchers.values())[0][1].batches
    binding.uninstall()
    server_hook._b200_encoders.clear()     # the module fixture owns (and closes) the encoder
    assert binding.stats["prefetched"] == len(clips) and tower.calls == 0
    for c, o in zip(clips, outs):
        alone, _ = enc.encode_pcm([c])
        torch.cuda.synchronize()
        assert torch.equal(o, alone)
    assert sum(batches) == len(clips) and len(batches) <= 3, f"the windows were not batched: {batches}"
