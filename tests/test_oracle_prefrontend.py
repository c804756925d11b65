"""Pins oracle/prefrontend.py (the CPU restatement of the reference's WS pre-frontend, src/server.py:26-42, 1321-1338)
against the library functions the reference itself calls (scipy) -- CPU only."""

import numpy as np
import pytest
import scipy.signal as ss

from oracle import prefrontend as pf


def _pcm16(n, seed, scale=6000.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = scale * (0.4 * rng.standard_normal(n) + np.sin(2 * np.pi * 440 * t) + 0.5 * np.sin(2 * np.pi * 50 * t) + 0.3)
    return np.clip(x, -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("n", [1, 7, 480, 9600, 40001])
def test_bandpass_restatement_is_bit_exact_vs_scipy_sosfilt(n):
    # the reference's call, verbatim (src/server.py:28-29)
    audio = _pcm16(n, n).astype(np.float32) / 32768.0
    sos = ss.butter(4, [300, 3400], btype="bandpass", fs=16000, output="sos")
    want = ss.sosfilt(sos, audio).astype(np.float32)
    got = pf.telephony_bandpass(audio)
    assert got.dtype == np.float32 and np.array_equal(got, want)


@pytest.mark.parametrize("up,down", [(2, 1), (1, 2), (160, 441), (3, 2), (1, 3), (320, 441)])
def test_resample_restatement_vs_scipy_resample_poly(up, down):
    x = _pcm16(6000, up * 1000 + down).astype(np.float64)
    want = ss.resample_poly(x, up, down)
    got = pf.resample_poly_f64(x, up, down)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    taps, half_len = pf.resample_taps(up, down)
    ref = ss.firwin(2 * half_len + 1, 1.0 / max(up, down), window=("kaiser", 5.0)) * up
    assert np.abs(taps - ref).max() <= 2e-15 * np.abs(ref).max()   # a few ulp: I0 series vs scipy.special.i0


def test_resample_pcm16_truncates_toward_zero_like_numpy_astype():
    x = _pcm16(4000, 5)
    y = pf.resample_pcm16(x, 8000)
    want = ss.resample_poly(x.astype(np.float64), 2, 1)
    assert y.dtype == np.int16 and y.shape[0] == 8000
    # identical except where the float64 sums straddle an integer by rounding noise (none expected on this input)
    assert np.array_equal(y, np.trunc(want).astype(np.int16))
    assert np.array_equal(pf.resample_pcm16(x, 16000), x)


def test_ws_window_layout():
    x = _pcm16(3000, 9)
    w = pf.ws_window(x, 16000, pad_silence=True)
    assert w.dtype == np.float32 and w.shape[0] == 3000 + 9600
    short = pf.ws_window(x[:1000], 16000, pad_silence=False)
    assert short.shape[0] == pf.MIN_SAMPLES and not short[1000:].any()
    # the flush silence IS filtered (the reference appends the zeros before the band-pass): the ringing is not zero
    assert np.abs(w[3000:3100]).max() > 0
    raw = pf.ws_window(x, 16000, bandpass=False, min_samples=0)
    assert np.array_equal(raw, x.astype(np.float32) / 32768.0)


# ---- upload normalisation (row a2): torchaudio sinc_interp_hann, the definition src/debug_audio.py:24-33 spells out
RATES = [8000, 11025, 22050, 32000, 44100, 48000]


def _wave(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.3 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 523.0 * t) * np.sin(2 * np.pi * 3 * t)).astype(np.float32)


@pytest.mark.parametrize("sr", RATES)
def test_sinc_hann_restatement_vs_torchaudio(sr):
    import torch
    import torchaudio.functional as AF

    x = _wave(sr // 2 + 37, sr)
    want = AF.resample(torch.from_numpy(x)[None], sr, 16000)[0].numpy()
    got = pf.resample_sinc_hann(x, sr, 16000)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() <= 1e-6, np.abs(got - want).max()


@pytest.mark.parametrize("sr", RATES)
def test_builtin_resample_taps_match_torchaudios_kernel(sr):
    """qasr_resample_f32_taps (pure host code of the library): the float32 recipe of _get_sinc_resample_kernel, <= 2 ulp per tap."""
    import ctypes as C
    import math

    import torch
    from torchaudio.functional.functional import _get_sinc_resample_kernel

    from qwen3_asr_b200 import load_library

    lib = load_library()
    k, w = _get_sinc_resample_kernel(sr, 16000, math.gcd(sr, 16000), dtype=torch.float32)
    k = k.numpy()[:, 0, :]
    n_ph, n_taps, width = C.c_int(), C.c_int(), C.c_int()
    assert lib.qasr_resample_f32_taps(sr, 16000, None, 0, C.byref(n_ph), C.byref(n_taps), C.byref(width)) == 0
    assert (n_ph.value, n_taps.value, width.value) == (k.shape[0], k.shape[1], w)
    buf = np.zeros(k.shape, np.float32)
    assert lib.qasr_resample_f32_taps(sr, 16000, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size, C.byref(n_ph), C.byref(n_taps),
                                      C.byref(width)) == 0
    assert np.abs(buf - k).max() <= 2.5e-7 * np.abs(k).max()
    ko, wo = pf.sinc_hann_kernel(sr, 16000)
    assert wo == w and np.abs(ko - k).max() <= 2.5e-7 * np.abs(k).max()
    assert lib.qasr_resample_f32_len(sr // 2 + 37, sr, 16000) == -(-(16000 * (sr // 2 + 37)) // sr)


def test_normalize_upload_is_debug_audio_py():
    import torch
    import torchaudio.functional as AF

    rng = np.random.default_rng(3)
    stereo = rng.uniform(-1, 1, size=(30011, 2))           # soundfile hands back float64 [frames, channels]
    mono = stereo.mean(axis=1)
    want = AF.resample(torch.from_numpy(mono).unsqueeze(0).float(), 44100, 16000).squeeze(0).numpy()   # debug_audio.py:24-33 verbatim
    got = pf.normalize_upload(stereo, 44100)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6


# ---- the SDK's long-audio splitter (8f-4), restated from the reference's description (LEARNING_LOG.md:215-219): parity unpinned
def _long_signal(seconds, seed, sr=16000, pauses=()):
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    x = (0.2 * rng.standard_normal(n) * (0.6 + 0.4 * np.sin(2 * np.pi * 0.7 * np.arange(n) / sr))).astype(np.float32)
    for t0, dur, level in pauses:
        a, b = int(t0 * sr), int((t0 + dur) * sr)
        x[a:b] = (level * rng.standard_normal(b - a)).astype(np.float32)
    return x


def test_split_points_cut_in_the_quietest_window():
    sr = 16000
    x = _long_signal(130, 1, pauses=[(38.0, 0.4, 1e-4), (80.0, 0.3, 1e-5), (97.0, 0.5, 1e-4)])
    b = pf.split_points(x, sr, max_chunk_sec=40.0, search_expand_sec=5.0, min_window_ms=100.0)
    assert b[0] == 0 and b[-1] == len(x) and np.all(np.diff(b) > 0)
    # first cut wanted at 40 s: the pause at 38.0-38.4 s lies inside [35, 45] -> the cut falls into it; the next cut is wanted 40 s later
    assert 38.0 * sr <= b[1] < 38.4 * sr
    assert abs(b[2] - (b[1] + 40 * sr)) <= 5 * sr and 80.0 * sr <= b[2] < 80.3 * sr
    assert np.all(np.diff(b) <= 45 * sr)
    # the float32 sliding-window formulation (np.convolve of |x| with a box, then the quietest sample of the winning window) picks
    # the same cuts on a signal whose pauses are unambiguous
    start, cuts = 0, []
    win, max_len, expand = int(0.1 * sr), 40 * sr, 5 * sr
    while len(x) - start > max_len:
        cut = start + max_len
        left, right = max(start, cut - expand), min(len(x), cut + expand)
        seg = np.abs(x[left:right])
        ws = np.convolve(seg, np.ones(win, dtype=np.float32), mode="valid")
        p = int(np.argmin(ws))
        cuts.append(left + p + int(np.argmin(seg[p:p + win])))
        start = cuts[-1]
    assert [abs(int(a) - int(c)) < win for a, c in zip(b[1:-1], cuts)] == [True] * len(cuts)
    # short audio: one chunk; all-zero audio: ties go to the first window / first sample
    assert pf.split_points(x[: 30 * sr], sr, 40.0).tolist() == [0, 30 * sr]
    z = pf.split_points(np.zeros(100 * sr, np.float32), sr, 40.0)
    assert z.tolist() == [0, 35 * sr, 70 * sr, 100 * sr]
