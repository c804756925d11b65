"""One process, several GPUs: host-side mirror of ``qasr_pool_*`` (include/qasr_b200.h).

Clips (SDK silence-split segments, SSE chunks, WebSocket windows) are independent, so N B200s are N replicas of the
encoder handle; a batch is cut into contiguous clip ranges of near-equal work and every GPU encodes its range on its
own stream in its own worker thread.  No collective, no NCCL.  Same calling convention as
``B200AudioEncoder.submit_pcm_host`` / ``wait``.
"""

from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import QasrError, check, load_library
from .encoder import _DTYPES, make_config, sinusoid_table


class B200EncoderPool:
    def __init__(self, cfg, weights: Mapping[str, "torch.Tensor | np.ndarray"], devices: Sequence[int] | None = None,
                 max_chunks: int = 0, max_tokens: int = 0, quantize: str | None = None, sharding: str = "auto"):
        if not torch.cuda.is_available():
            raise QasrError("B200EncoderPool needs CUDA devices; this backend has no CPU path")
        self.lib = load_library()
        self.devices = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
        self.cfg = make_config(cfg, max_chunks, max_tokens, quantize)
        self.output_dim = int(self.cfg.output_dim)
        devs = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(self.lib, self.lib.qasr_pool_create(C.byref(self.cfg), devs, len(self.devices), C.byref(h)), "qasr_pool_create")
        self._h = h
        try:
            modes = {"auto": 0, "contiguous": 1, "lpt": 2}
            if sharding not in modes:
                raise QasrError(f"unknown sharding mode {sharding!r}; expected one of {sorted(modes)}")
            check(self.lib, self.lib.qasr_pool_set_sharding(self._h, modes[sharding]), "qasr_pool_set_sharding")
            for name, w in weights.items():
                if "positional_embedding" in name and name != "positional_embedding":
                    continue
                self._set_weight(name, w)
            if "positional_embedding" not in weights:
                self._set_weight("positional_embedding", sinusoid_table(13, int(self.cfg.d_model)))
            check(self.lib, self.lib.qasr_pool_finalize(self._h), "qasr_pool_finalize")
        except Exception:
            self.close()
            raise

    def _set_weight(self, name: str, w) -> None:
        t = torch.as_tensor(w) if not isinstance(w, torch.Tensor) else w
        t = t.detach()
        if t.dtype not in _DTYPES:
            t = t.float()
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        check(self.lib, self.lib.qasr_pool_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), _DTYPES[t.dtype], shape, t.dim()),
              f"qasr_pool_set_weight({name})")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.qasr_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.lib.qasr_pool_size(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.qasr_pool_workspace_bytes(self._h))

    def submit_pcm_host(self, pcm_host: torch.Tensor, offsets: np.ndarray, out_host: torch.Tensor):
        """Enqueue one batch across the pool; returns (ticket, token_lens, device of every clip).  The offsets array is
        kept alive by the returned ticket object's owner: keep ``offsets`` referenced until ``collect``."""
        assert not pcm_host.is_cuda and pcm_host.dtype == torch.float32 and pcm_host.is_contiguous()
        assert not out_host.is_cuda and out_host.dtype == torch.bfloat16 and out_host.is_contiguous()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        toks = np.zeros(n, dtype=np.int64)
        devs = np.zeros(n, dtype=np.int32)
        ticket = C.c_uint64(0)
        check(self.lib, self.lib.qasr_pool_submit(self._h, C.c_void_p(pcm_host.data_ptr()), offsets.ctypes.data_as(_lib._I64P), n,
                                                  C.c_void_p(out_host.data_ptr()), int(out_host.shape[0]), toks.ctypes.data_as(_lib._I64P),
                                                  devs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ticket)),
              "qasr_pool_submit")
        return int(ticket.value), toks, devs

    def stats(self, reset: bool = False) -> list:
        """Per-member averages: where a shard's time goes (host enqueue / host wait / device H2D, compute, D2H / turnaround), ms."""
        n = len(self)
        buf = (C.c_double * (7 * n))()
        check(self.lib, self.lib.qasr_pool_stats(self._h, buf, n, 1 if reset else 0), "qasr_pool_stats")
        names = ("shards", "enqueue_ms", "wait_ms", "h2d_ms", "compute_ms", "d2h_ms", "turnaround_ms")
        return [{"device": self.devices[i], **{k: round(buf[7 * i + j], 4) for j, k in enumerate(names)}} for i in range(n)]

    def collect(self, ticket: int) -> None:
        check(self.lib, self.lib.qasr_pool_collect(self._h, C.c_uint64(int(ticket))), "qasr_pool_collect")

    def encode_pcm(self, clips: Sequence["np.ndarray | torch.Tensor"]):
        """List of 16 kHz float32 clips -> (bf16 hidden states on the host [sum tokens, output_dim], token_lens)."""
        offs = np.zeros(len(clips) + 1, dtype=np.int64)
        for i, c in enumerate(clips):
            offs[i + 1] = offs[i] + int(c.shape[0])
        pcm = torch.empty(int(offs[-1]), dtype=torch.float32, pin_memory=True)
        for i, c in enumerate(clips):
            pcm[int(offs[i]):int(offs[i + 1])] = torch.as_tensor(c, dtype=torch.float32)
        total = int(sum(int(self.lib.qasr_token_len(int(c.shape[0]) // 160)) for c in clips))
        out = torch.empty((total, self.output_dim), dtype=torch.bfloat16, pin_memory=True)
        ticket, toks, _ = self.submit_pcm_host(pcm, offs, out)
        self.collect(ticket)
        return out, toks
