"""WS pre-frontend on the GPU (through the C ABI) vs the scipy-pinned oracle: int16 -> [resample] -> /32768 -> band-pass."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pre():
    from oracle import CONFIGS, make_weights
    from qwen3_asr_b200 import B200AudioEncoder, B200PreFrontend

    cfg = CONFIGS["tiny"]
    enc = B200AudioEncoder(cfg, make_weights(cfg, seed=1), max_chunks=64)
    yield B200PreFrontend(enc)
    enc.close()


def _pcm16(n, seed, scale=6000.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = scale * (0.4 * rng.standard_normal(n) + np.sin(2 * np.pi * 440 * t) + 0.5 * np.sin(2 * np.pi * 50 * t) + 0.3)
    return np.clip(x, -32768, 32767).astype(np.int16)


ABS_FLOOR = 1e-15   # the blocked evaluation forgets state below 1e-18 of the signal scale: where the true output has itself
                    # decayed to that level (the tail of the flush silence) only the absolute error is meaningful


def _ulp_diff(a, b):
    """float32 ulps between a and b; differences below ABS_FLOOR count as 0."""
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    d = np.abs(ai - bi)
    return np.where(np.abs(a.astype(np.float64) - b.astype(np.float64)) <= ABS_FLOOR, 0, d)


def test_bandpass_batch_matches_scipy_bitwise(pre):
    import scipy.signal as ss

    # ragged batch incl. lengths below / at / above the 512-sample thread block and the warm-up span, and an empty stream
    lens = [1, 511, 512, 513, 1568, 5000, 0, 96000, 33333]
    wins = [_pcm16(n, 100 + i) for i, n in enumerate(lens)]
    pre.min_samples = 0
    out, offs = pre.prepare(wins, 16000, pad_silence=None, bandpass=True)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    sos = ss.butter(4, [300, 3400], btype="bandpass", fs=16000, output="sos")
    assert offs.tolist() == np.concatenate([[0], np.cumsum(lens)]).tolist()
    n_diff = 0
    for i, w in enumerate(wins):
        if w.shape[0] == 0:
            assert offs[i + 1] == offs[i]       # empty window in, empty clip out (the reference returns early, server.py:1331)
            continue
        want = ss.sosfilt(sos, w.astype(np.float32) / 32768.0).astype(np.float32)   # the reference's own two lines
        got = out[offs[i]:offs[i + 1]]
        d = _ulp_diff(got, want)
        assert d.max(initial=0) <= 1, (i, int(d.max()))       # float64 cascade: only a rounding tie could move a float32 by one ulp
        n_diff += int((d > 0).sum())
    assert n_diff <= 2, n_diff
    pre.min_samples = 8000


def test_ws_windows_c3_shape_vs_oracle(pre):
    from oracle import prefrontend as opf

    # BASELINE config 3 shape: ragged windows, every 4th one flushing (+600 ms of silence), short ones padded to 0.5 s
    lens = [min(96000, 7200 * (1 + (5 * i) % 14)) for i in range(16)] + [3000]
    wins = [_pcm16(n, 200 + i) for i, n in enumerate(lens)]
    flush = [i % 4 == 3 for i in range(len(wins))]
    out, offs = pre.prepare(wins, 16000, pad_silence=flush, bandpass=True)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i, w in enumerate(wins):
        want = opf.ws_window(w, 16000, pad_silence=flush[i])
        got = out[offs[i]:offs[i + 1]]
        assert got.shape == want.shape, i
        assert _ulp_diff(got, want).max() <= 1, i
    assert offs[-1] - offs[-2] == 8000 and not out[offs[-2] + 3000:offs[-1]].any()


@pytest.mark.parametrize("orig_sr", [8000, 44100, 48000, 22050])
def test_resample_matches_oracle_bitwise(pre, orig_sr):
    from oracle import prefrontend as opf

    wins = [_pcm16(n, 300 + i) for i, n in enumerate([4000, 1, 777, 12345])]
    dev, offs = pre._pack_int16(wins)
    out, ooffs = pre.resample_pcm16_packed(dev, offs, orig_sr)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    up, down = opf.resample_ratio(orig_sr)
    taps, _ = opf.resample_taps(up, down)
    for i, w in enumerate(wins):
        want = opf.resample_pcm16(w, orig_sr)
        got = out[ooffs[i]:ooffs[i + 1]]
        assert got.shape == want.shape
        # built-in taps come from the C library's libm, the oracle's from numpy: allow the int16 truncation to flip on < 0.1 % of samples
        assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1
        assert (got != want).mean() <= 1e-3
    # with the oracle's own taps handed over the ABI the float64 sums are the same operations in the same order: bit-exact
    out2, ooffs2 = pre.resample_pcm16_packed(dev, offs, orig_sr, taps=taps)
    torch.cuda.synchronize()
    out2 = out2.cpu().numpy()
    for i, w in enumerate(wins):
        assert np.array_equal(out2[ooffs2[i]:ooffs2[i + 1]], opf.resample_pcm16(w, orig_sr)), i


def test_reference_named_helpers(pre):
    from oracle import prefrontend as opf

    x = _pcm16(2400, 4)
    assert pre.resample_pcm_bytes(x.tobytes(), 16000) == x.tobytes()            # unchanged, as in the reference
    y = np.frombuffer(pre.resample_pcm_bytes(x.tobytes(), 8000), dtype=np.int16)
    assert y.shape[0] == 4800 and np.abs(y.astype(np.int32) - opf.resample_pcm16(x, 8000)).max() <= 1
    f = pre.telephony_bandpass(x)
    assert f.dtype == np.float32 and _ulp_diff(f, opf.telephony_bandpass(x.astype(np.float32) / 32768.0)).max() <= 1


def test_pcm_bytes_to_tokens_stays_on_gpu_and_matches_two_step_path(pre):
    from oracle import prefrontend as opf

    wins = [_pcm16(n, 400 + i) for i, n in enumerate([16000, 48000, 7000])]
    flush = [False, True, False]
    hid, toks = pre.encode_windows([w.tobytes() for w in wins], 16000, pad_silence=flush)
    clips = [opf.ws_window(w, 16000, pad_silence=f) for w, f in zip(wins, flush)]
    hid2, toks2 = pre.enc.encode_pcm(clips)
    torch.cuda.synchronize()
    assert toks.tolist() == toks2.tolist()
    # identical PCM (<= 1 ulp on a vanishing fraction of samples) -> identical bf16 tokens except where a rounding boundary is hit
    a, b = hid.float().cpu().numpy(), hid2.float().cpu().numpy()
    assert np.abs(a - b).max() <= 2e-2 * max(1.0, np.abs(b).max())
    assert (a != b).mean() < 0.02


def test_rejects_bad_arguments(pre):
    from qwen3_asr_b200 import QasrError

    with pytest.raises(QasrError):
        pre._pack_int16([np.zeros(10, np.float32)])
    dev, offs = pre._pack_int16([_pcm16(100, 1)])
    import ctypes as C
    from qwen3_asr_b200 import _lib
    out = torch.empty(100, dtype=torch.float32, device=dev.device)
    ooffs = np.zeros(2, dtype=np.int64)
    bad_sos = np.array([[1.0, 0, 0, 2.0, 0, 0]])      # a0 != 1: scipy.signal.sosfilt raises, so do we
    rc = pre.lib.qasr_ws_window(pre.enc._h, C.c_void_p(dev.data_ptr()), offs.ctypes.data_as(_lib._I64P), 1, None,
                                bad_sos.ctypes.data_as(C.POINTER(C.c_double)), 1, 0, C.c_void_p(out.data_ptr()), 100,
                                ooffs.ctypes.data_as(_lib._I64P), None)
    assert rc != 0 and b"all ones" in pre.lib.qasr_last_error()
    rc = pre.lib.qasr_ws_window(pre.enc._h, C.c_void_p(dev.data_ptr()), offs.ctypes.data_as(_lib._I64P), 1, None, None, 0, 0,
                                C.c_void_p(out.data_ptr()), 50, ooffs.ctypes.data_as(_lib._I64P), None)
    assert rc != 0 and b"out_capacity" in pre.lib.qasr_last_error()


def test_window_batcher_over_the_real_encoder(pre):
    """concurrently submitted windows are gathered into one ragged encoder call; every client gets exactly the
    tokens a lone call would have produced (the kernels are batch-invariant)."""
    import threading

    from qwen3_asr_b200.batcher import WindowBatcher

    wins = [_pcm16(n, 500 + i).tobytes() for i, n in enumerate([16000, 9000, 48000, 20000, 31234, 8000])]
    flush = [i % 3 == 2 for i in range(len(wins))]
    b = WindowBatcher(lambda w, f: pre.encode_windows(w, 16000, pad_silence=f), max_wait_ms=200.0)
    futs = [None] * len(wins)

    def client(i):
        futs[i] = b.submit(wins[i], flush=flush[i])
    ts = [threading.Thread(target=client, args=(i,)) for i in range(len(wins))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    outs = [f.result(timeout=30) for f in futs]
    b.close()
    assert sum(b.batches) == len(wins) and len(b.batches) <= 2
    for i in range(len(wins)):
        alone, _ = pre.encode_windows([wins[i]], 16000, pad_silence=[flush[i]])
        torch.cuda.synchronize()
        assert torch.equal(outs[i], alone), i


# ---- row a2: upload normalisation (any rate, any channel count -> mono float32 16 kHz) -----------------------------------
@pytest.mark.parametrize("sr", [8000, 22050, 44100, 48000, 11025, 16000])
def test_resample_f32_vs_torchaudio(pre, sr):
    """within 1e-6 absolute of torchaudio.functional.resample (sinc_interp_hann, width 6, rolloff 0.99 -- the
    definition src/debug_audio.py:28-33 uses), ragged batch, built-in taps and torchaudio's own kernel handed over the ABI."""
    import math

    import torchaudio.functional as AF
    from torchaudio.functional.functional import _get_sinc_resample_kernel

    rng = np.random.default_rng(sr)
    lens = [sr * 2 + 37, 1, sr // 3, 5 * sr + 1234, 0, 441]
    clips = [(0.5 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 700.0 * np.arange(n) / sr)).astype(np.float32) for n in lens]
    out, offs = pre.normalize_audio(clips, sr)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i, c in enumerate(clips):
        want = AF.resample(torch.from_numpy(c)[None], sr, 16000)[0].numpy() if len(c) else np.zeros(0, np.float32)
        got = out[offs[i]:offs[i + 1]]
        assert got.shape == want.shape, (i, got.shape, want.shape)
        if len(want):
            assert np.abs(got - want).max() <= 1e-6, (sr, i, np.abs(got - want).max())
    if sr != 16000:
        k, _ = _get_sinc_resample_kernel(sr, 16000, math.gcd(sr, 16000), dtype=torch.float32)
        out2, offs2 = pre.normalize_audio(clips, sr, taps=k.numpy()[:, 0, :])
        torch.cuda.synchronize()
        assert offs2.tolist() == offs.tolist()
        want = np.concatenate([AF.resample(torch.from_numpy(c)[None], sr, 16000)[0].numpy() for c in clips if len(c)])
        assert np.abs(out2.cpu().numpy() - want).max() <= 5e-7     # same taps: only the summation order differs


def test_stereo_upload_is_debug_audio_py(pre):
    """src/debug_audio.py:24-33 verbatim on the CPU (float64 mean over channels, .float(), torchaudio resample) vs the device path."""
    import torchaudio.functional as AF

    from oracle import prefrontend as opf

    rng = np.random.default_rng(9)
    clips = [rng.uniform(-1, 1, size=(n, 2)) for n in (44100, 30011)]
    out, offs = pre.normalize_audio(clips, 44100)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i, a in enumerate(clips):
        mono = a.astype(np.float32).astype(np.float64).mean(axis=1)     # the device receives float32 samples
        want = AF.resample(torch.from_numpy(mono).unsqueeze(0).float(), 44100, 16000).squeeze(0).numpy()
        got = out[offs[i]:offs[i + 1]]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6
        assert np.abs(got - opf.normalize_upload(a.astype(np.float32), 44100)).max() <= 1e-6
    # 16 kHz stereo: only the channel mean
    out, offs = pre.normalize_audio([clips[0]], 16000)
    want = clips[0].astype(np.float32).astype(np.float64).mean(axis=1).astype(np.float32)
    assert np.array_equal(out.cpu().numpy(), want)


def test_upload_to_tokens_stays_on_the_device(pre):
    """44.1 kHz stereo upload -> normalise -> log-mel -> encoder equals encoding the torchaudio-normalised clip."""
    import torchaudio.functional as AF

    rng = np.random.default_rng(10)
    t = np.arange(3 * 44100) / 44100.0
    a = np.stack([0.4 * np.sin(2 * np.pi * 330 * t) + 0.05 * rng.standard_normal(len(t)), 0.3 * np.sin(2 * np.pi * 550 * t)], axis=1)
    hid, toks = pre.encode_uploads([a], 44100)
    mono = a.astype(np.float32).astype(np.float64).mean(axis=1)
    ref_pcm = AF.resample(torch.from_numpy(mono).unsqueeze(0).float(), 44100, 16000).squeeze(0).numpy()
    hid2, toks2 = pre.enc.encode_pcm([ref_pcm])
    torch.cuda.synchronize()
    assert toks.tolist() == toks2.tolist() == [39]
    a_, b_ = hid.float().cpu().numpy(), hid2.float().cpu().numpy()
    # the two PCM inputs differ by <= 1e-6 on every sample: bf16 roundings flip here and there, nothing more
    assert np.abs(a_ - b_).max() <= 2e-2 * max(1.0, np.abs(b_).max())


# ---- 8f-4: the SDK's long-audio splitter on the device ----------------------------------------------------------------------
def test_split_audio_matches_the_oracle_bit_for_bit(pre):
    from oracle import prefrontend as opf
    from test_oracle_prefrontend import _long_signal

    sr = 16000
    cases = [
        (_long_signal(130, 1, pauses=[(38.0, 0.4, 1e-4), (80.0, 0.3, 1e-5), (97.0, 0.5, 1e-4)]), 40.0, 5.0, 100.0),
        (_long_signal(75, 2), 20.0, 3.0, 50.0),                       # no pauses: the minimum is wherever the noise dips
        (np.zeros(100 * sr, np.float32), 40.0, 5.0, 100.0),           # all ties: first window, first sample
        (_long_signal(61, 3, pauses=[(59.9, 1.1, 0.0)]), 30.0, 5.0, 100.0),   # search range clipped at the end of the audio
        (_long_signal(30, 4), 40.0, 5.0, 100.0),                      # shorter than one chunk
    ]
    for x, max_s, exp_s, win_ms in cases:
        dev = torch.from_numpy(x).to(pre.enc.tdev)
        got = pre.split_audio(dev, max_chunk_sec=max_s, search_expand_sec=exp_s, min_window_ms=win_ms)
        want = opf.split_points(x, sr, max_s, exp_s, win_ms)
        assert got.tolist() == want.tolist()


def test_split_chunks_feed_the_encoder_as_clip_offsets(pre):
    """boundaries = clip_offsets: the chunks of a long recording are encoded as independent clips straight from the same buffer."""
    from test_oracle_prefrontend import _long_signal

    x = _long_signal(50, 7, pauses=[(14.0, 0.5, 1e-4), (31.0, 0.5, 1e-4)])
    dev = torch.from_numpy(x).to(pre.enc.tdev)
    b = pre.split_audio(dev, max_chunk_sec=15.0, search_expand_sec=3.0)
    assert len(b) == 5
    hid, toks = pre.enc.encode_pcm_packed(dev, b)
    parts = [pre.enc.encode_pcm([x[b[i]:b[i + 1]]])[0] for i in range(len(b) - 1)]
    torch.cuda.synchronize()
    assert torch.equal(hid, torch.cat(parts))
