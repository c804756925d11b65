"""WebSocket pre-frontend oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates what the reference does to a WS audio window before the log-mel:

* ``int16_to_float``      <- src/server.py:1335-1336  ``np.frombuffer(.., int16).astype(float32) / 32768.0``
* ``resample_pcm16``      <- src/server.py:32-42 ``_resample_pcm_bytes``: int16 -> float -> resample -> ``astype(int16)``
                             (numpy's float->int cast truncates toward zero)
* ``telephony_bandpass``  <- src/server.py:26-29 ``_telephony_bandpass``: ``sosfilt(butter(4, [300, 3400], "bandpass",
                             fs=16000, output="sos"), audio).astype(float32)`` -- scipy computes the cascade in float64
* ``ws_window``           <- src/server.py:1321-1338 ``_transcribe_with_context``: [overlap + chunk] (+ 600 ms of int16 zeros
                             when flushing) -> /32768 -> band-pass, plus the SDK's 0.5 s minimum-length zero pad

Parity pinning
--------------
* The band-pass is pinned: ``sosfilt_df2t`` restates scipy's ``_sosfilt`` (direct form II transposed, float64) and
  tests/test_oracle_prefrontend.py checks it bit-exactly against ``scipy.signal.sosfilt`` (scipy 1.18.1, the very function the
  reference calls).
* The resampler is **parity unpinned against the reference**: the reference calls ``librosa.resample`` (default
  ``res_type="soxr_hq"``); neither librosa nor soxr is installed here and there is no network.  The definition used
  instead is the polyphase Kaiser-windowed-sinc of ``scipy.signal.resample_poly`` (published algorithm, restated in
  ``resample_poly_f64``), and that restatement IS pinned against scipy's own output.
* The upload normalisation (``normalize_upload`` = mono mean + ``resample_sinc_hann``) restates the reference's own spelling of
  that step, src/debug_audio.py:24-33 (``audio.mean(axis=1)`` then ``torchaudio.functional.resample``: sinc_interp_hann,
  lowpass_filter_width 6, rolloff 0.99 -- torchaudio/functional/functional.py:1305-1400, 1403-1430).  It is **pinned** against
  the in-container torchaudio 2.11 (tests/test_oracle_prefrontend.py); the SDK's internal resampler behind src/server.py:867
  stays unavailable offline.
"""

from __future__ import annotations

import math

import numpy as np

TARGET_SR = 16000
WS_FLUSH_SILENCE_MS = 600   # src/config.py / .env.example default
MIN_SAMPLES = 8000          # SDK 0.5 s minimum


def int16_to_float(pcm16: np.ndarray) -> np.ndarray:
    return pcm16.astype(np.float32) / np.float32(32768.0)


# ---------------------------------------------------------------------------------------------------------------
# resample_poly (scipy/signal/_signaltools.py resample_poly + _fir_filter_design.py firwin + windows.kaiser)
# ---------------------------------------------------------------------------------------------------------------
def _bessel_i0(x: np.ndarray) -> np.ndarray:
    # power series, converged to double precision for the |x| <= 5 used here
    x = np.asarray(x, dtype=np.float64)
    q = (x / 2.0) ** 2
    term = np.ones_like(x)
    s = np.ones_like(x)
    for k in range(1, 64):
        term = term * q / (k * k)
        s = s + term
    return s


def resample_taps(up: int, down: int, beta: float = 5.0):
    """(taps * up, half_len) of scipy's default design: firwin(2*half_len+1, 1/max_rate, window=('kaiser', 5.0))."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    m = np.arange(n, dtype=np.float64) - 0.5 * (n - 1)
    cutoff = 1.0 / max_rate
    h = cutoff * np.sinc(cutoff * m)
    alpha = (n - 1) / 2.0
    win = _bessel_i0(beta * np.sqrt(np.maximum(0.0, 1.0 - ((np.arange(n) - alpha) / alpha) ** 2))) / _bessel_i0(np.float64(beta))
    h = h * win
    h = h / h.sum()   # unit gain at DC (firwin scale=True, pass_zero)
    return h * up, half_len


def resample_ratio(orig_sr: int, target_sr: int = TARGET_SR):
    g = math.gcd(int(orig_sr), int(target_sr))
    return int(target_sr) // g, int(orig_sr) // g


def resample_poly_f64(x: np.ndarray, up: int, down: int) -> np.ndarray:
    """y[n] = sum_j h[j] x_up[n*down + half_len - j], x_up = x with up-1 zeros between samples; len = ceil(N*up/down)."""
    x = np.asarray(x, dtype=np.float64)
    if up == down:
        return x.copy()
    h, half_len = resample_taps(up, down)
    n_in = x.shape[0]
    n_out = -(-n_in * up // down)
    y = np.zeros(n_out, dtype=np.float64)
    n = np.arange(n_out, dtype=np.int64)
    m0 = n * down + half_len
    # taps of phase (m0 mod up) in ascending j: the summation order the CUDA kernel uses as well
    j = m0 % up
    for _ in range((len(h) + up - 1) // up):
        i = (m0 - j) // up
        ok = (j < len(h)) & (i >= 0) & (i < n_in)
        y = y + np.where(ok, h[np.minimum(j, len(h) - 1)] * x[np.clip(i, 0, n_in - 1)], 0.0)
        j = j + up
    return y


def resample_pcm16(pcm16: np.ndarray, orig_sr: int, target_sr: int = TARGET_SR) -> np.ndarray:
    """_resample_pcm_bytes on samples: int16 in, int16 out (truncation toward zero, saturating instead of numpy's wrap)."""
    if orig_sr == target_sr:
        return pcm16.copy()
    up, down = resample_ratio(orig_sr, target_sr)
    y = resample_poly_f64(pcm16.astype(np.float64), up, down)
    return np.clip(np.trunc(y), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------------------------------
# band-pass (scipy/signal/_sosfilt.pyx _sosfilt_float: direct form II transposed per section, sample-major)
# ---------------------------------------------------------------------------------------------------------------
def telephony_sos(sr: int = TARGET_SR) -> np.ndarray:
    from scipy.signal import butter

    return butter(4, [300, 3400], btype="bandpass", fs=sr, output="sos")


def sosfilt_df2t(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """float64 cascade, zero initial state:  y = b0 x + s0;  s0 = b1 x - a1 y + s1;  s1 = b2 x - a2 y  (per section)."""
    sos = np.asarray(sos, dtype=np.float64)
    cur = np.asarray(x, dtype=np.float64).copy()
    for b0, b1, b2, a0, a1, a2 in sos:
        b0, b1, b2, a1, a2 = b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0
        s0 = 0.0
        s1 = 0.0
        out = np.empty_like(cur)
        for n, xn in enumerate(cur.tolist()):
            yn = b0 * xn + s0
            s0 = b1 * xn - a1 * yn + s1
            s1 = b2 * xn - a2 * yn
            out[n] = yn
        cur = out
    return cur


def telephony_bandpass(audio: np.ndarray, sr: int = TARGET_SR) -> np.ndarray:
    return sosfilt_df2t(telephony_sos(sr), audio).astype(np.float32)


def ws_window(pcm16: np.ndarray, orig_sr: int = TARGET_SR, pad_silence: bool = False, bandpass: bool = True,
              min_samples: int = MIN_SAMPLES) -> np.ndarray:
    """One WS window: int16 samples at orig_sr -> float32 16 kHz clip ready for the log-mel."""
    x16 = resample_pcm16(np.asarray(pcm16, dtype=np.int16), orig_sr)
    if pad_silence:
        x16 = np.concatenate([x16, np.zeros(int(WS_FLUSH_SILENCE_MS / 1000 * TARGET_SR), np.int16)])
    f = int16_to_float(x16)
    if bandpass:
        f = telephony_bandpass(f)
    if f.shape[0] < min_samples:
        f = np.concatenate([f, np.zeros(min_samples - f.shape[0], np.float32)])
    return f


# ---------------------------------------------------------------------------------------------------------------
# torchaudio.functional.resample (sinc_interp_hann), restated in numpy float32 where torchaudio computes in float32
def sinc_hann_kernel(orig_sr: int, new_sr: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """_get_sinc_resample_kernel for a float32 waveform (functional.py:1340-1400): every step in float32, in torch's order.
    Returns (kernel float32 [new, 2 * width + orig], width) with orig / new reduced by their gcd."""
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx
    t = t * f32(base_freq)
    t = np.clip(t, f32(-lowpass_filter_width), f32(lowpass_filter_width))
    window = np.cos(t * f32(math.pi) / f32(lowpass_filter_width) / f32(2)) ** 2
    t = t * f32(math.pi)
    scale = f32(base_freq / orig)
    with np.errstate(invalid="ignore", divide="ignore"):
        kernels = np.where(t == 0, f32(1.0), np.sin(t) / t).astype(f32)
    kernels = kernels * (window * scale)
    return kernels.astype(f32), width


def resample_sinc_hann(x: np.ndarray, orig_sr: int, new_sr: int) -> np.ndarray:
    """_apply_sinc_resample_kernel (functional.py:1403-1430): zero-pad (width, width + orig), correlate with stride orig, one phase
    per output residue, cut to ceil(new * n / orig).  float32 taps and samples, products summed in float64."""
    x = np.asarray(x, dtype=np.float32).reshape(-1)
    if int(orig_sr) == int(new_sr):
        return x.copy()
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    kern, width = sinc_hann_kernel(orig_sr, new_sr)
    n = x.shape[0]
    xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + orig, np.float32)]).astype(np.float64)
    n_taps = kern.shape[1]
    n_blocks = (xp.shape[0] - n_taps) // orig + 1
    frames = np.lib.stride_tricks.sliding_window_view(xp, n_taps)[::orig][:n_blocks]      # [blocks, taps]
    y = frames @ kern.astype(np.float64).T                                                  # [blocks, new]
    target = -(-new * n // orig)
    return y.reshape(-1)[:target].astype(np.float32)


def normalize_upload(audio: np.ndarray, sr: int, target_sr: int = TARGET_SR) -> np.ndarray:
    """src/debug_audio.py:24-33: (frames,) or (frames, channels) array as soundfile returns it -> mono float32 at target_sr."""
    a = np.asarray(audio)
    if a.ndim > 1:
        a = a.astype(np.float64).mean(axis=1)
    return resample_sinc_hann(a.astype(np.float32), sr, target_sr)


# ---------------------------------------------------------------------------------------------------------------
# The SDK's long-audio splitter (split_audio_into_chunks), as the reference describes it: LEARNING_LOG.md:215-219 "sliding window
# convolution with +/-5s search range", chunks of at most 1200 s (CLAUDE.md "up to 20min").  The SDK source is not available
# offline, so this is a restatement of that description -- PARITY UNPINNED against the SDK.  Exact integer arithmetic defines the
# result bit for bit: |x| -> floor(|x| * 2^40) as int64, window sums by differences of an int64 prefix sum, first minimum wins.
def split_points(wav: np.ndarray, sr: int = TARGET_SR, max_chunk_sec: float = 1200.0, search_expand_sec: float = 5.0,
                 min_window_ms: float = 100.0) -> np.ndarray:
    x = np.asarray(wav, dtype=np.float32).reshape(-1)
    n = x.shape[0]
    max_len, expand = int(max_chunk_sec * sr), int(search_expand_sec * sr)
    win = max(4, int((min_window_ms / 1000.0) * sr))
    bounds, start = [0], 0
    while n - start > max_len:
        cut = start + max_len
        left, right = max(start, cut - expand), min(n, cut + expand)
        boundary = cut
        if right - left > win:
            q = np.floor(np.abs(x[left:right].astype(np.float64)) * float(2 ** 40)).astype(np.int64)
            ps = np.concatenate([[0], np.cumsum(q)])
            sums = ps[win:] - ps[:-win]
            wstart = int(np.argmin(sums))
            boundary = left + wstart + int(np.argmin(q[wstart:wstart + win]))
        boundary = min(max(boundary, start + 1), n)
        bounds.append(boundary)
        start = boundary
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)
