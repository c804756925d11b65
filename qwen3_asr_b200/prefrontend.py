"""Host-side mirror of the reference's WebSocket pre-frontend, backed by libqasr_b200.so.

The reference prepares every WS window on the CPU before the SDK sees it:

* ``_resample_pcm_bytes(pcm_bytes, orig_sr)``  -- src/server.py:32-42, per incoming message
* ``_telephony_bandpass(audio, sr)``           -- src/server.py:26-29
* the prologue of ``_transcribe_with_context``  -- src/server.py:1321-1338: [overlap + chunk] (+ 600 ms flush
  silence) -> int16 -> float32 / 32768 -> band-pass
* the normalisation of an HTTP upload (src/server.py:867 hands ``(audio, sr)`` to the SDK): mono mean + resample to 16 kHz,
  in the reference's own spelling of it, src/debug_audio.py:24-33 (``audio.mean(axis=1)``, ``torchaudio.functional.resample``)

``B200PreFrontend`` keeps those names and argument meanings, but takes a *batch* of streams and leaves the result on
the GPU, packed exactly as ``B200AudioEncoder.encode_pcm_packed`` / ``logmel_packed`` expect it, so PCM bytes in ->
encoder tokens out never touches the CPU between the socket and the decoder.  No CPU fallback: a missing library or GPU raises.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import QasrError, check

TARGET_SR = 16000
WS_FLUSH_SILENCE_MS = 600      # reference default (src/config.py, .env.example)
MIN_CLIP_SAMPLES = 8000        # the SDK's 0.5 s minimum clip length


def telephony_sos(sr: int = TARGET_SR) -> np.ndarray:
    """The reference's filter design, verbatim call: butter(4, [300, 3400], btype="bandpass", fs=sr, output="sos")
    (src/server.py:28).  scipy is a dependency of the reference server itself; only the 4 x 6 coefficients cross the C ABI."""
    from scipy.signal import butter

    return np.ascontiguousarray(butter(4, [300, 3400], btype="bandpass", fs=sr, output="sos"), dtype=np.float64)


def resample_ratio(orig_sr: int, target_sr: int = TARGET_SR):
    g = math.gcd(int(orig_sr), int(target_sr))
    return int(target_sr) // g, int(orig_sr) // g


class B200PreFrontend:
    def __init__(self, encoder, min_samples: int = MIN_CLIP_SAMPLES):
        self.enc = encoder            # B200AudioEncoder: owns the handle, the device and the stream convention
        self.lib = encoder.lib
        self.min_samples = int(min_samples)
        self._sos = {}

    # ---- helpers ------------------------------------------------------------------------------------------
    def _pack_int16(self, streams: Sequence["np.ndarray | bytes | bytearray | torch.Tensor"]):
        arrs = []
        for s in streams:
            if isinstance(s, (bytes, bytearray, memoryview)):
                s = np.frombuffer(s, dtype=np.int16)          # raw PCM 16-bit LE, as the socket delivers it
            elif isinstance(s, torch.Tensor):
                s = s.cpu().numpy()
            s = np.asarray(s)
            if s.dtype != np.int16:
                raise QasrError(f"pre-frontend input must be int16 PCM, got {s.dtype}")
            arrs.append(s.reshape(-1))
        offs = np.zeros(len(arrs) + 1, dtype=np.int64)
        for i, a in enumerate(arrs):
            offs[i + 1] = offs[i] + a.shape[0]
        host = torch.empty(int(offs[-1]), dtype=torch.int16, pin_memory=True)
        for i, a in enumerate(arrs):
            host[int(offs[i]):int(offs[i + 1])] = torch.from_numpy(np.ascontiguousarray(a))
        return host.to(self.enc.tdev, non_blocking=True), offs

    # ---- _resample_pcm_bytes ----------------------------------------------------------------------------------
    def resample_pcm16_packed(self, pcm16: torch.Tensor, offsets: np.ndarray, orig_sr: int, target_sr: int = TARGET_SR,
                              taps: np.ndarray | None = None):
        """int16 device tensor of streams back to back at orig_sr -> (int16 device tensor at target_sr, offsets)."""
        assert pcm16.is_cuda and pcm16.dtype == torch.int16 and pcm16.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if orig_sr == target_sr:
            return pcm16, offsets
        up, down = resample_ratio(orig_sr, target_sr)
        n = len(offsets) - 1
        total = int(sum(self.lib.qasr_resample_len(int(offsets[i + 1] - offsets[i]), up, down) for i in range(n)))
        out = torch.empty(total, dtype=torch.int16, device=self.enc.tdev)
        out_offs = np.zeros(n + 1, dtype=np.int64)
        tp, nt = None, 0
        if taps is not None:
            taps = np.ascontiguousarray(taps, dtype=np.float64)
            tp, nt = taps.ctypes.data_as(C.POINTER(C.c_double)), int(taps.shape[0])
        check(self.lib, self.lib.qasr_resample_pcm16(self.enc._h, C.c_void_p(pcm16.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                     up, down, tp, nt, C.c_void_p(out.data_ptr()), total,
                                                     out_offs.ctypes.data_as(_lib._I64P), self.enc._stream()),
              "qasr_resample_pcm16")
        return out, out_offs

    def resample_pcm_bytes(self, pcm_bytes: bytes, orig_sr: int) -> bytes:
        """Same signature and return type as the reference's _resample_pcm_bytes (one message)."""
        if orig_sr == TARGET_SR:
            return pcm_bytes
        dev, offs = self._pack_int16([pcm_bytes])
        out, _ = self.resample_pcm16_packed(dev, offs, orig_sr)
        return out.cpu().numpy().tobytes()

    # ---- upload normalisation: any rate, any channel count -> mono float32 16 kHz -----------------------------------
    def normalize_audio(self, clips: Sequence["np.ndarray | torch.Tensor"], sr: int, target_sr: int = TARGET_SR,
                        taps: np.ndarray | None = None):
        """clips: arrays shaped (frames,) or (frames, channels) -- what ``soundfile.read`` returns -- all at ``sr`` and with the
        same channel count.  Returns (mono float32 device tensor of clips back to back at target_sr, offsets): ready for
        ``encode_pcm_packed`` / ``logmel_packed``.  Channel mean and sinc_interp_hann resampling run on the device."""
        arrs = [np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c) for c in clips]
        channels = {1 if a.ndim == 1 else int(a.shape[1]) for a in arrs}
        if len(channels) > 1:
            raise QasrError(f"normalize_audio: clips of one call must share a channel count, got {sorted(channels)}")
        ch = channels.pop() if channels else 1
        offs = np.zeros(len(arrs) + 1, dtype=np.int64)
        for i, a in enumerate(arrs):
            offs[i + 1] = offs[i] + a.shape[0]
        host = torch.empty(int(offs[-1]) * ch, dtype=torch.float32, pin_memory=True)
        for i, a in enumerate(arrs):
            host[int(offs[i]) * ch:int(offs[i + 1]) * ch] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1))
        return self.resample_f32_packed(host.to(self.enc.tdev, non_blocking=True), offs, ch, sr, target_sr, taps)

    def resample_f32_packed(self, pcm: torch.Tensor, offsets: np.ndarray, channels: int, orig_sr: int, target_sr: int = TARGET_SR,
                            taps: np.ndarray | None = None):
        """pcm: float32 device tensor, frames of ``channels`` interleaved samples, streams back to back; offsets in frames."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if channels == 1 and int(orig_sr) == int(target_sr):
            return pcm, offsets
        total = int(sum(self.lib.qasr_resample_f32_len(int(offsets[i + 1] - offsets[i]), int(orig_sr), int(target_sr)) for i in range(n)))
        out = torch.empty(total, dtype=torch.float32, device=self.enc.tdev)
        out_offs = np.zeros(n + 1, dtype=np.int64)
        tp = None
        if taps is not None:
            taps = np.ascontiguousarray(taps, dtype=np.float32)
            tp = taps.ctypes.data_as(C.POINTER(C.c_float))
        check(self.lib, self.lib.qasr_resample_f32(self.enc._h, C.c_void_p(pcm.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                   int(channels), int(orig_sr), int(target_sr), tp, C.c_void_p(out.data_ptr()), total,
                                                   out_offs.ctypes.data_as(_lib._I64P), self.enc._stream()),
              "qasr_resample_f32")
        return out, out_offs

    # ---- the SDK's long-audio splitter -----------------------------------------------------------------------------
    def split_audio(self, pcm: torch.Tensor, max_chunk_sec: float = 1200.0, search_expand_sec: float = 5.0, min_window_ms: float = 100.0,
                    sr: int = TARGET_SR) -> np.ndarray:
        """mono float32 device tensor -> int64 boundaries [n_chunks + 1] (0 ... len): chunks of at most ~max_chunk_sec cut at the
        quietest point within +/- search_expand_sec of every nominal cut (LEARNING_LOG.md:215-219).  The result is the
        ``clip_offsets`` of ``encode_pcm_packed`` / ``B200EncoderPool.submit_pcm_host`` for the same buffer."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous() and pcm.dim() == 1
        n = int(pcm.shape[0])
        cap = n // max(1, int(max_chunk_sec * sr) - int(search_expand_sec * sr)) + 4
        out = np.zeros(cap, dtype=np.int64)
        nch = C.c_int(0)
        check(self.lib, self.lib.qasr_split_audio(self.enc._h, C.c_void_p(pcm.data_ptr()), n, int(sr), float(max_chunk_sec),
                                                  float(search_expand_sec), float(min_window_ms), out.ctypes.data_as(_lib._I64P), cap,
                                                  C.byref(nch), self.enc._stream()),
              "qasr_split_audio")
        return out[: nch.value + 1].copy()

    def encode_uploads(self, clips: Sequence, sr: int):
        """(audio, sr) uploads -> (bf16 hidden states, token_lens): normalise, log-mel and encode without leaving the GPU."""
        pcm, offs = self.normalize_audio(clips, sr)
        return self.enc.encode_pcm_packed(pcm, offs)

    # ---- int16 -> float -> band-pass ------------------------------------------------------------------------------
    def ws_window_packed(self, pcm16: torch.Tensor, offsets: np.ndarray, pad_silence: "Sequence[bool] | None" = None,
                         bandpass: bool = True, sr: int = TARGET_SR):
        """int16 16 kHz windows on the device -> (float32 device tensor of clips back to back, offsets): ready for
        ``encode_pcm_packed``.  pad_silence[i]: append WS_FLUSH_SILENCE_MS of zeros before filtering (the flush path)."""
        assert pcm16.is_cuda and pcm16.dtype == torch.int16 and pcm16.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        pads = np.zeros(n, dtype=np.int32)
        if pad_silence is not None:
            silence = int((WS_FLUSH_SILENCE_MS / 1000) * sr)
            pads[:] = [silence if p else 0 for p in pad_silence]
        lens = (offsets[1:] - offsets[:-1]) + pads
        total = int(np.where(lens > 0, np.maximum(lens, self.min_samples), 0).sum())
        out = torch.empty(total, dtype=torch.float32, device=self.enc.tdev)
        out_offs = np.zeros(n + 1, dtype=np.int64)
        sos_p, ns = None, 0
        if bandpass:
            if sr not in self._sos:
                self._sos[sr] = telephony_sos(sr)
            sos = self._sos[sr]
            sos_p, ns = sos.ctypes.data_as(C.POINTER(C.c_double)), int(sos.shape[0])
        check(self.lib, self.lib.qasr_ws_window(self.enc._h, C.c_void_p(pcm16.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                pads.ctypes.data_as(C.POINTER(C.c_int32)), sos_p, ns, self.min_samples,
                                                C.c_void_p(out.data_ptr()), total, out_offs.ctypes.data_as(_lib._I64P),
                                                self.enc._stream()),
              "qasr_ws_window")
        return out, out_offs

    def telephony_bandpass(self, audio_int16: np.ndarray, sr: int = TARGET_SR) -> np.ndarray:
        """_telephony_bandpass for one int16 window (the reference feeds it int16 / 32768): float32 numpy out."""
        dev, offs = self._pack_int16([audio_int16])
        keep, self.min_samples = self.min_samples, 0
        try:
            out, _ = self.ws_window_packed(dev, offs, None, True, sr)
        finally:
            self.min_samples = keep
        return out.cpu().numpy()

    # ---- the whole WS prologue, batched --------------------------------------------------------------------------
    def prepare(self, windows: Sequence, orig_sr: int = TARGET_SR, pad_silence: "Sequence[bool] | None" = None,
                bandpass: bool = True):
        """windows: int16 arrays / PCM byte strings at orig_sr, one per stream.  Returns (pcm float32 on the device,
        clip offsets).  Note: the reference resamples per incoming *message* and concatenates afterwards; pass messages
        through ``resample_pcm16_packed`` individually if that edge behaviour matters."""
        dev, offs = self._pack_int16(windows)
        if orig_sr != TARGET_SR:
            dev, offs = self.resample_pcm16_packed(dev, offs, orig_sr)
        return self.ws_window_packed(dev, offs, pad_silence, bandpass)

    def encode_windows(self, windows: Sequence, orig_sr: int = TARGET_SR, pad_silence: "Sequence[bool] | None" = None,
                       bandpass: bool = True):
        """PCM bytes of N concurrent WS windows -> (bf16 hidden states [sum tokens, output_dim], token_lens)."""
        pcm, offs = self.prepare(windows, orig_sr, pad_silence, bandpass)
        return self.enc.encode_pcm_packed(pcm, offs)
