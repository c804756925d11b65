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
