"""Host-side mirror of the reference's encoder interface, backed by libqasr_b200.so.

``B200AudioEncoder`` stands where the reference's two optional encoder backends stand
(``_trt_encoder`` / ``_onnx_session``, /root/reference/src/server.py:237-251, 461-475, 873-914): an
object created once at model load that turns ``input_features`` into hidden states.  Its
``forward`` keeps the signature of the module it replaces --
``Qwen3OmniMoeAudioEncoder.forward(input_features, feature_lens=..., aftercnn_lens=...)``
(transformers modeling_qwen3_omni_moe.py:698-766) -- and ``logmel`` replaces
``WhisperFeatureExtractor._torch_extract_fbank_features`` (feature_extraction_whisper.py:135-164).

PyTorch is used for tensor hand-off only (allocation, streams); every FLOP runs in the library.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Mapping, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import QASR_BF16, QASR_F16, QASR_F32, QasrConfig, QasrError, check, load_library

HOP = 160
N_MELS = 128

_CFG_FIELDS = (
    "d_model", "encoder_layers", "encoder_attention_heads", "encoder_ffn_dim", "output_dim",
    "n_window", "n_window_infer", "downsample_hidden_size", "num_mel_bins", "max_source_positions",
)


@dataclass
class EncoderOutput:
    """Shaped like transformers' BaseModelOutput for the one field the SDK reads."""

    last_hidden_state: torch.Tensor

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


def _cfg_value(cfg, name: str, aliases: Sequence[str] = ()):
    for n in (name, *aliases):
        if isinstance(cfg, Mapping) and n in cfg:
            return int(cfg[n])
        if hasattr(cfg, n):
            return int(getattr(cfg, n))
    raise KeyError(f"audio config has no field {name!r}")


QASR_FLAG_FP8, QASR_FLAG_FP8_PER_ROW = 1, 2
_QUANT_FLAGS = {None: 0, "": 0, "bf16": 0, "fp8": QASR_FLAG_FP8, "fp8_per_tensor": QASR_FLAG_FP8,
                "fp8_per_row": QASR_FLAG_FP8 | QASR_FLAG_FP8_PER_ROW}


def make_config(cfg, max_chunks: int = 0, max_tokens: int = 0, quantize: str | None = None) -> QasrConfig:
    """From a HF ``audio_config`` (or a dict / the oracle's EncoderConfig) to the C struct.
    ``quantize``: None / "fp8" (per-tensor scales, the reference's QUANTIZE=fp8, src/server.py:362-371) / "fp8_per_row"."""
    alias = {
        "encoder_layers": ("layers",), "encoder_attention_heads": ("heads",), "encoder_ffn_dim": ("ffn",),
        "downsample_hidden_size": ("downsample_hidden",),
    }
    vals = {n: _cfg_value(cfg, n, alias.get(n, ())) for n in _CFG_FIELDS}
    if quantize not in _QUANT_FLAGS:
        raise QasrError(f"unknown quantize mode {quantize!r}; expected one of {sorted(k for k in _QUANT_FLAGS if k)}")
    return QasrConfig(**vals, max_chunks=max_chunks, max_tokens=max_tokens, flags=_QUANT_FLAGS[quantize])


def sinusoid_table(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """SinusoidsPositionEmbedding (modeling_qwen3_omni_moe.py:88-106), computed the same way in torch."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2).float())
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


_DTYPES = {torch.float32: QASR_F32, torch.bfloat16: QASR_BF16, torch.float16: QASR_F16}


class B200AudioEncoder:
    def __init__(self, cfg, weights: Mapping[str, "torch.Tensor | np.ndarray"], device: int = 0,
                 max_chunks: int = 0, max_tokens: int = 0, quantize: str | None = None):
        if not torch.cuda.is_available():
            raise QasrError("B200AudioEncoder needs a CUDA device; this backend has no CPU path")
        self.lib = load_library()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self.cfg = make_config(cfg, max_chunks, max_tokens, quantize)
        self.quantize = quantize
        self.output_dim = int(self.cfg.output_dim)
        self.d_model = int(self.cfg.d_model)
        h = C.c_void_p()
        check(self.lib, self.lib.qasr_create(C.byref(self.cfg), self.device, C.byref(h)), "qasr_create")
        self._h = h
        self.weight_bytes = 0   # bf16 bytes of the parameters handed over: what one forward must at least read from HBM
        try:
            for name, w in weights.items():
                if "positional_embedding" in name and name != "positional_embedding":
                    continue
                self._set_weight(name, w)
                self.weight_bytes += 2 * int(np.prod(tuple(w.shape)))
            if "positional_embedding" not in weights:
                self._set_weight("positional_embedding", sinusoid_table(13, self.d_model))
            check(self.lib, self.lib.qasr_finalize(self._h), "qasr_finalize")
        except Exception:
            self.close()
            raise

    # ---- construction helpers -----------------------------------------------------------------
    @classmethod
    def from_module(cls, audio_tower, device: int | None = None, **kw) -> "B200AudioEncoder":
        """From the live ``m.model.thinker.audio_tower`` torch module (weights are copied)."""
        sd = audio_tower.state_dict()
        if device is None:
            p = next(audio_tower.parameters())
            device = p.device.index if p.is_cuda else 0
        return cls(audio_tower.config, sd, device=device, **kw)

    def _set_weight(self, name: str, w) -> None:
        t = torch.as_tensor(w) if not isinstance(w, torch.Tensor) else w
        t = t.detach()
        if t.dtype not in _DTYPES:
            t = t.float()
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        check(self.lib, self.lib.qasr_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), _DTYPES[t.dtype], shape, t.dim()),
              f"qasr_set_weight({name})")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.qasr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def token_len(self, feature_len: int) -> int:
        return int(self.lib.qasr_token_len(int(feature_len)))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.qasr_workspace_bytes(self._h))

    def pack_clips(self, clips: Sequence["np.ndarray | torch.Tensor"]):
        """List of mono 16 kHz float32 clips -> (packed device tensor, int64 offsets array)."""
        offs = np.zeros(len(clips) + 1, dtype=np.int64)
        for i, c in enumerate(clips):
            offs[i + 1] = offs[i] + int(c.shape[0])
        host = torch.empty(int(offs[-1]), dtype=torch.float32, pin_memory=True)
        for i, c in enumerate(clips):
            host[int(offs[i]):int(offs[i + 1])] = torch.as_tensor(c, dtype=torch.float32)
        return host.to(self.tdev, non_blocking=True), offs

    # ---- the operators ------------------------------------------------------------------------
    def logmel_packed(self, pcm: torch.Tensor, offsets: np.ndarray):
        """pcm: float32 device tensor of clips back to back; offsets int64 [n+1].
        Returns (mel float32 [128, sum T] on the device, feature_lens int64 array)."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        flens = np.zeros(n, dtype=np.int64)
        cols = int(sum((int(offsets[i + 1]) - int(offsets[i])) // HOP for i in range(n)))
        mel = torch.empty((N_MELS, cols), dtype=torch.float32, device=self.tdev)
        check(self.lib, self.lib.qasr_logmel(self._h, C.c_void_p(pcm.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                             C.c_void_p(mel.data_ptr()), cols, flens.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_logmel")
        return mel, flens

    def logmel(self, clips: Sequence["np.ndarray | torch.Tensor"]):
        pcm, offs = self.pack_clips(clips)
        return self.logmel_packed(pcm, offs)

    def encode(self, mel: torch.Tensor, feature_lens) -> torch.Tensor:
        """mel: [128, >= sum T] float32 or bfloat16 on the device (packed clips); returns bf16 [sum tokens, output_dim]."""
        assert mel.is_cuda and mel.dim() == 2 and mel.shape[0] == N_MELS and mel.stride(1) == 1
        if mel.dtype not in (torch.float32, torch.bfloat16):
            mel = mel.float()
        flens = np.ascontiguousarray(np.asarray(torch.as_tensor(feature_lens).cpu()), dtype=np.int64).reshape(-1)
        n = int(flens.shape[0])
        toks = np.zeros(n, dtype=np.int64)
        total = int(sum(self.token_len(int(t)) for t in flens))
        out = torch.empty((total, self.output_dim), dtype=torch.bfloat16, device=self.tdev)
        check(self.lib, self.lib.qasr_encode(self._h, C.c_void_p(mel.data_ptr()), _DTYPES[mel.dtype], int(mel.stride(0)),
                                             flens.ctypes.data_as(_lib._I64P), n, C.c_void_p(out.data_ptr()),
                                             toks.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_encode")
        self.last_token_lens = toks
        return out

    def encode_into(self, mel: torch.Tensor, feature_lens, inputs_embeds: torch.Tensor, audio_mask: torch.Tensor):
        """Encode and write every audio token straight into its placeholder row of the decoder's ``inputs_embeds`` -- the fused
        form of ``inputs_embeds.masked_scatter(audio_mask[..., None], audio_features)`` (modeling_qwen3_omni_moe.py:2135-2143).
        ``inputs_embeds``: bf16 [..., hidden] on this device with hidden == output_dim; ``audio_mask``: bool [...] (one flag per
        position, true at the audio placeholders, in the order the clips were given).  Returns the token lengths."""
        assert mel.is_cuda and mel.dim() == 2 and mel.shape[0] == N_MELS and mel.stride(1) == 1
        if mel.dtype not in (torch.float32, torch.bfloat16):
            mel = mel.float()
        assert inputs_embeds.is_cuda and inputs_embeds.dtype == torch.bfloat16 and inputs_embeds.is_contiguous()
        assert inputs_embeds.shape[-1] == self.output_dim and tuple(audio_mask.shape) == tuple(inputs_embeds.shape[:-1])
        flens = np.ascontiguousarray(np.asarray(torch.as_tensor(feature_lens).cpu()), dtype=np.int64).reshape(-1)
        n = int(flens.shape[0])
        total = int(sum(self.token_len(int(t)) for t in flens))
        rows = torch.nonzero(audio_mask.reshape(-1).to(self.tdev), as_tuple=False).reshape(-1).to(torch.int64).contiguous()
        if int(rows.numel()) != total:
            raise QasrError(f"audio_mask marks {int(rows.numel())} positions but the clips produce {total} audio tokens")
        toks = np.zeros(n, dtype=np.int64)
        check(self.lib, self.lib.qasr_encode_scatter(self._h, C.c_void_p(mel.data_ptr()), _DTYPES[mel.dtype], int(mel.stride(0)),
                                                     flens.ctypes.data_as(_lib._I64P), n, C.c_void_p(inputs_embeds.data_ptr()),
                                                     int(inputs_embeds.shape[-1]), C.c_void_p(rows.data_ptr()),
                                                     toks.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_encode_scatter")
        self._keep_rows = rows   # the kernels read it asynchronously on the stream
        return toks

    def encode_pcm_packed(self, pcm: torch.Tensor, offsets: np.ndarray, out: torch.Tensor | None = None):
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        toks = np.zeros(n, dtype=np.int64)
        total = int(sum(self.token_len((int(offsets[i + 1]) - int(offsets[i])) // HOP) for i in range(n)))
        if out is None:
            out = torch.empty((total, self.output_dim), dtype=torch.bfloat16, device=self.tdev)
        assert out.is_cuda and out.dtype == torch.bfloat16 and out.shape[0] >= total and out.is_contiguous()
        check(self.lib, self.lib.qasr_encode_pcm(self._h, C.c_void_p(pcm.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                 C.c_void_p(out.data_ptr()), toks.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_encode_pcm")
        return out[:total], toks

    def encode_pcm(self, clips: Sequence["np.ndarray | torch.Tensor"]):
        """PCM clips -> (bf16 hidden states [sum tokens, output_dim], token_lens)."""
        pcm, offs = self.pack_clips(clips)
        return self.encode_pcm_packed(pcm, offs)

    def encode_pcm_host(self, pcm_host: torch.Tensor, offsets: np.ndarray, out_host: torch.Tensor):
        """End to end with HOST buffers (pinned recommended): H2D, log-mel, encoder, D2H, sync."""
        assert not pcm_host.is_cuda and pcm_host.dtype == torch.float32 and pcm_host.is_contiguous()
        assert not out_host.is_cuda and out_host.dtype == torch.bfloat16 and out_host.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        toks = np.zeros(n, dtype=np.int64)
        check(self.lib, self.lib.qasr_encode_pcm_host(self._h, C.c_void_p(pcm_host.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                      C.c_void_p(out_host.data_ptr()), int(out_host.shape[0]),
                                                      toks.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_encode_pcm_host")
        return toks

    def submit_pcm_host(self, pcm_host: torch.Tensor, offsets: np.ndarray, out_host: torch.Tensor):
        """Pipelined end-to-end call: enqueue H2D + log-mel + encoder + D2H and return (ticket, token_lens) at once;
        ``wait(ticket)`` blocks until ``out_host`` is filled.  Keep at most two tickets un-waited."""
        assert not pcm_host.is_cuda and pcm_host.dtype == torch.float32 and pcm_host.is_contiguous()
        assert not out_host.is_cuda and out_host.dtype == torch.bfloat16 and out_host.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        toks = np.zeros(n, dtype=np.int64)
        ticket = C.c_uint64(0)
        check(self.lib, self.lib.qasr_submit_pcm_host(self._h, C.c_void_p(pcm_host.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                      C.c_void_p(out_host.data_ptr()), int(out_host.shape[0]),
                                                      toks.ctypes.data_as(_lib._I64P), self._stream(), C.byref(ticket)),
              "qasr_submit_pcm_host")
        return int(ticket.value), toks

    def wait(self, ticket: int) -> None:
        check(self.lib, self.lib.qasr_wait(self._h, C.c_uint64(int(ticket))), "qasr_wait")

    def poll(self, ticket: int) -> bool:
        """Non-blocking: has ``out_host`` of ``ticket`` been filled?  (``wait`` would return at once.)"""
        done = C.c_int(0)
        check(self.lib, self.lib.qasr_poll(self._h, C.c_uint64(int(ticket)), C.byref(done)), "qasr_poll")
        return bool(done.value)

    def logmel_host(self, clips: Sequence[np.ndarray]):
        offs = np.zeros(len(clips) + 1, dtype=np.int64)
        for i, c in enumerate(clips):
            offs[i + 1] = offs[i] + int(c.shape[0])
        pcm = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32) for c in clips]))
        flens = np.zeros(len(clips), dtype=np.int64)
        cols = int(sum(int(c.shape[0]) // HOP for c in clips))
        mel = np.empty((N_MELS, cols), dtype=np.float32)
        check(self.lib, self.lib.qasr_logmel_host(self._h, pcm.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(_lib._I64P), len(clips),
                                                  mel.ctypes.data_as(C.c_void_p), flens.ctypes.data_as(_lib._I64P), self._stream()),
              "qasr_logmel_host")
        return mel, flens

    # ---- drop-in forward (the callable the server hook installs) -----------------------------------
    def forward(self, input_features=None, feature_lens=None, aftercnn_lens=None, **kwargs):
        """Same contract as Qwen3OmniMoeAudioEncoder.forward: packed ``input_features`` [128, sum T]
        plus ``feature_lens``; also accepts the reference hook's single padded clip [1, 128, T]
        (src/server.py:876-882) in which case the whole T is one clip."""
        x = input_features
        if x is None:
            raise QasrError("forward called without input_features")
        if x.dim() == 3:
            if feature_lens is None:
                feature_lens = [x.shape[2]] * x.shape[0]
            lens = [int(v) for v in torch.as_tensor(feature_lens).reshape(-1).tolist()]
            x = torch.cat([x[i, :, : lens[i]] for i in range(x.shape[0])], dim=1)
        elif feature_lens is None:
            feature_lens = [x.shape[1]]
        if not x.is_cuda:
            x = x.to(self.tdev)
        if x.stride(1) != 1:
            x = x.contiguous()
        out = self.encode(x, feature_lens)
        return EncoderOutput(last_hidden_state=out)

    __call__ = forward

    # ---- measurement ---------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self.lib.qasr_launch_count(self._h))

    def graph_stats(self) -> dict:
        """{"graphs": live CUDA graphs of repeated call shapes, "replays": calls served by one, "bytes": device bytes they hold}"""
        n, r, b = C.c_int(0), C.c_uint64(0), C.c_size_t(0)
        check(self.lib, self.lib.qasr_graph_stats(self._h, C.byref(n), C.byref(r), C.byref(b)), "qasr_graph_stats")
        return {"graphs": int(n.value), "replays": int(r.value), "bytes": int(b.value)}

    def profile(self, on: bool = True) -> None:
        """Bracket every kernel launch with CUDA events on the launching stream (clears old records)."""
        check(self.lib, self.lib.qasr_profile_enable(self._h, 1 if on else 0), "qasr_profile_enable")

    def profile_read(self) -> dict:
        """{kernel name: {"ms": summed ms, "work": summed algorithmic FLOPs (bytes for logmel), "launches": n}}"""
        cap = 64
        names = C.create_string_buffer(4096)
        ms = (C.c_double * cap)()
        work = (C.c_double * cap)()
        cnt = (C.c_int32 * cap)()
        n = C.c_int(0)
        check(self.lib, self.lib.qasr_profile_read(self._h, names, 4096, ms, work, cnt, cap, C.byref(n)), "qasr_profile_read")
        keys = names.value.decode().split("\n")[: n.value]
        return {k: {"ms": ms[i], "work": work[i], "launches": int(cnt[i])} for i, k in enumerate(keys)}

    # ---- bring-up hooks ------------------------------------------------------------------------
    def debug_read(self, name: str, dtype=np.uint16, max_bytes: int = 1 << 30) -> np.ndarray:
        buf = np.empty(max_bytes, dtype=np.uint8)
        n = C.c_size_t(0)
        check(self.lib, self.lib.qasr_debug_read(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), max_bytes, C.byref(n)),
              f"qasr_debug_read({name})")
        return buf[: n.value].view(dtype).copy()


def bf16_bits_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)
