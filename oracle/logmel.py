"""Log-mel oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, per clip and with standalone semantics:

* ``mel_filter_bank``  <- transformers/audio_utils.py:263-332 (hertz_to_mel,
  mel_to_hertz, slaney), :356-375 (_create_triangular_filter_bank), :453-544
  (mel_filter_bank, norm="slaney"), as called at
  models/whisper/feature_extraction_whisper.py:95-103.
* ``logmel``           <- feature_extraction_whisper.py:135-164
  (_torch_extract_fbank_features): hann(400) periodic, hop 160, centred reflect
  pad, drop last frame, |X|^2, mel, log10(clamp 1e-10), max(x, max-8), (x+4)/4.
* ``feature_len``      <- feature_extraction_whisper.py:328-337 for one clip.

``logmel`` computes in float64 (the torch-f32 path it restates carries ~2e-5 of
its own rounding noise); ``logmel_torch_f32`` is the same
arithmetic through ``torch.stft`` in float32, i.e. what the reference's CPU
process actually runs, and is the one timed as the CPU baseline.
"""

from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
N_MELS = 128
SAMPLE_RATE = 16000


def _hertz_to_mel_slaney(freq):
    # audio_utils.py:263-295 (mel_scale="slaney")
    freq = np.asarray(freq, dtype=np.float64)
    min_log_hertz = 1000.0
    min_log_mel = 15.0
    logstep = 27.0 / np.log(6.4)
    mels = 3.0 * freq / 200.0
    log_region = freq >= min_log_hertz
    mels = np.where(log_region, min_log_mel + np.log(np.maximum(freq, 1e-30) / min_log_hertz) * logstep, mels)
    return mels


def _mel_to_hertz_slaney(mels):
    # audio_utils.py:298-332
    mels = np.asarray(mels, dtype=np.float64)
    min_log_hertz = 1000.0
    min_log_mel = 15.0
    logstep = np.log(6.4) / 27.0
    freq = 200.0 * mels / 3.0
    log_region = mels >= min_log_mel
    freq = np.where(log_region, min_log_hertz * np.exp(logstep * (mels - min_log_mel)), freq)
    return freq


def mel_filter_bank(n_freq: int = N_FFT // 2 + 1, n_mels: int = N_MELS, fmin: float = 0.0,
                    fmax: float = 8000.0, sr: int = SAMPLE_RATE) -> np.ndarray:
    """[n_freq, n_mels] float64 Slaney-scale, Slaney-normalised triangular bank."""
    mel_min = _hertz_to_mel_slaney(fmin)
    mel_max = _hertz_to_mel_slaney(fmax)
    mel_pts = np.linspace(mel_min, mel_max, n_mels + 2)
    filter_freqs = _mel_to_hertz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    # audio_utils.py:356-375
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(np.zeros(1), np.minimum(down, up))
    # norm="slaney" (audio_utils.py:533-536)
    enorm = 2.0 / (filter_freqs[2 : n_mels + 2] - filter_freqs[:n_mels])
    fb = fb * np.expand_dims(enorm, 0)
    return fb


def feature_len(n_samples: int) -> int:
    """Mel frames of one clip processed alone: floor(N / 160)."""
    return int(n_samples) // HOP


def _frames_f64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n <= N_FFT // 2:
        raise ValueError("reflect padding needs more than 200 samples (torch.stft raises too)")
    padded = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    t = n // HOP  # the last of the n//160 + 1 frames is dropped (:150)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(t)[:, None]
    return padded[idx]


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def log10_mel_unclamped(x: np.ndarray) -> np.ndarray:
    """[128, T] float64 log10(max(mel, 1e-10)) before the max-8 clamp and scaling."""
    frames = _frames_f64(x) * hann_periodic()[None, :]
    spec = np.fft.rfft(frames, axis=1)
    power = spec.real**2 + spec.imag**2  # [T, 201]
    fb32 = mel_filter_bank().astype(np.float32).astype(np.float64)  # f64 -> f32 cast at :152
    mel = power @ fb32  # [T, 128]
    return np.log10(np.maximum(mel, 1e-10)).T


def logmel(x: np.ndarray) -> np.ndarray:
    """Standalone per-clip log-mel, float64 arithmetic, returned as float32 [128, T]."""
    log_spec = log10_mel_unclamped(x)
    if log_spec.shape[1] == 0:
        return np.zeros((N_MELS, 0), dtype=np.float32)
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def logmel_torch_f32(x: np.ndarray, threads: int | None = None) -> np.ndarray:
    """Same arithmetic through torch.stft in float32 (the reference's CPU code path)."""
    import torch

    if threads is not None:
        torch.set_num_threads(threads)
    wav = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    window = torch.hann_window(N_FFT)
    stft = torch.stft(wav, N_FFT, HOP, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    fb = torch.from_numpy(mel_filter_bank()).to(torch.float32)
    mel = fb.T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).numpy()
