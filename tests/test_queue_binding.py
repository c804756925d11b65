"""The window batcher bound beneath the reference's PriorityInferQueue (qwen3_asr_b200/queue_binding.py).

CPU only: the backend and the batch encoder are recording fakes, the queue is the REFERENCE's own ``PriorityInferQueue``
(imported from /root/reference/src/server.py when that tree exists -- this container) or a line-for-line equivalent stand-in with
the same ``submit(fn, priority)`` contract where it does not (the GPU box).  The real CUDA encoder behind the same binding is
covered by tests/test_gpu_prefrontend.py::test_queue_binding_over_the_real_encoder."""

import asyncio
import concurrent.futures
import heapq
import os
import sys
import time
import types

import numpy as np
import pytest
import torch

from qwen3_asr_b200 import server_hook
from qwen3_asr_b200.queue_binding import QueueBinding, job_window

REF_SRC = "/root/reference/src"


class FakeTower(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.calls = 0

    def forward(self, input_features, feature_lens=None, aftercnn_lens=None):
        self.calls += 1
        return types.SimpleNamespace(last_hidden_state=torch.zeros(1, 4))


class FakeSDKModel:
    """.model.thinker.audio_tower + .transcribe((audio, sr)): computes T = N // 160 like the extractor and calls the tower once"""

    def __init__(self, delay=0.0):
        self.tower = FakeTower()
        self.model = types.SimpleNamespace(thinker=types.SimpleNamespace(audio_tower=self.tower))
        self.delay = delay

    def transcribe(self, audio_sr, language=None, return_time_stamps=False):
        audio, sr = audio_sr
        t = len(audio) // 160
        out = self.model.thinker.audio_tower.forward(torch.zeros(128, t), feature_lens=torch.tensor([t]))
        time.sleep(self.delay)     # the decoder
        return [types.SimpleNamespace(text=f"{float(out.last_hidden_state[0, 0]):.1f}", language="en")]


class FakeBackend:
    """per-call encode: marks its output with -1 so a test can tell it from a batched result"""

    def __init__(self):
        self.calls = 0

    def forward(self, input_features, feature_lens=None):
        self.calls += 1
        return types.SimpleNamespace(last_hidden_state=torch.full((1, 4), -1.0))

    def close(self):
        pass


class StandInQueue:
    """Same contract as src/server.py:59-107: heap by (priority, submit_time), one worker, one job at a time on one thread."""

    def __init__(self):
        self._heap, self._seq = [], 0
        self._has_work = asyncio.Event()
        self._executor = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        self._task = None

    def start(self):
        self._task = asyncio.create_task(self._worker())

    async def _worker(self):
        loop = asyncio.get_event_loop()
        while True:
            await self._has_work.wait()
            if not self._heap:
                self._has_work.clear()
                continue
            _, _, _, fut, fn = heapq.heappop(self._heap)
            if not self._heap:
                self._has_work.clear()
            try:
                fut.set_result(await loop.run_in_executor(self._executor, fn))
            except Exception as e:
                fut.set_exception(e)

    async def submit(self, fn, priority: int = 1):
        fut = asyncio.get_event_loop().create_future()
        self._seq += 1
        heapq.heappush(self._heap, (priority, time.time(), self._seq, fut, fn))
        self._has_work.set()
        return await fut


def _make_server(monkeypatch, fast=False):
    """A module object with the names the binding uses: the reference's server.py itself when importable."""
    server = None
    if os.path.isdir(REF_SRC):
        monkeypatch.setenv("MODEL_ID", "Qwen/Qwen3-ASR-1.7B")
        monkeypatch.syspath_prepend(REF_SRC)
        sys.modules.pop("server", None)
        try:
            import server as ref_server  # noqa: the reference module, unmodified

            server = ref_server
            monkeypatch.setattr(server, "_infer_queue", server.PriorityInferQueue(), raising=False)
            monkeypatch.setattr(server, "_cuda_stream", None, raising=False)
            monkeypatch.setattr(server, "_PINNED_AUDIO_BUFFER", None, raising=False)
            monkeypatch.setattr(server, "USE_SPECULATIVE", False, raising=False)
        except Exception:
            server = None
    if server is None:
        server = types.SimpleNamespace(_infer_queue=StandInQueue(), USE_SPECULATIVE=False)

        def _do_transcribe(audio, sr, lang_code, return_timestamps, use_fast=False):
            m = server._fast_model if (use_fast and server._fast_model is not None) else server.model
            return m.transcribe((audio, sr), language=lang_code, return_time_stamps=return_timestamps)

        server._do_transcribe = _do_transcribe
    monkeypatch.setattr(server, "model", FakeSDKModel(delay=0.02), raising=False)
    monkeypatch.setattr(server, "_fast_model", FakeSDKModel(delay=0.0) if fast else None, raising=False)
    return server


@pytest.fixture(autouse=True)
def _clean(monkeypatch):
    server_hook.unload()
    monkeypatch.setenv("B200_ENCODER", "1")
    yield
    server_hook.unload()
    sys.modules.pop("server", None)


def _recording_factory(record):
    """encode_factory for QueueBinding: every batch is recorded; window i of a batch gets hidden == its first sample * 1000"""

    def factory(backend):
        def encode(windows, _flush):
            record.append((backend, [len(w) for w in windows]))
            hidden = torch.cat([torch.full((1, 4), float(w[0]) * 1000.0) for w in windows])
            return hidden, [1] * len(windows)

        return encode

    return factory


def test_job_window_reads_the_three_call_sites():
    audio, sr, lang_code, pad_silence = np.ones(16000, np.float32), 16000, None, True
    ws = lambda: (audio, sr, lang_code, not pad_silence)                      # server.py:1349-1355
    assert job_window(ws)[1:] == (16000, False) and job_window(ws)[0] is audio
    pad_silence = False
    ws2 = lambda: (audio, sr, lang_code, not pad_silence)
    assert job_window(ws2)[2] is True
    chunk = audio[100:900]
    sse = lambda c=chunk: (c, sr, lang_code)                                  # server.py:985-988
    assert job_window(sse)[0] is chunk
    assert job_window(lambda: 1) is None and job_window(print) is None


def test_concurrent_submits_become_one_encode(monkeypatch):
    """N WebSocket windows submitted concurrently (the C3 workload): ONE batched encode serves all of them, every job still
    goes through the reference's queue and _do_transcribe one at a time, and each gets its own window's result."""
    server = _make_server(monkeypatch)
    backend = FakeBackend()
    server_hook.install(server, batch_windows=False)
    assert server_hook.try_load_b200_encoder(server.model, factory=lambda t: backend) == 1
    record = []
    binding = QueueBinding(server, max_wait_ms=20.0, encode_factory=_recording_factory(record))
    binding.install()

    async def ws_window(i):
        audio = np.full(16000 + 160 * i, (i + 1) / 1000.0, dtype=np.float32)
        sr, lang_code, pad_silence = 16000, None, True
        res = await server._infer_queue.submit(lambda: server._do_transcribe(audio, sr, lang_code, False, use_fast=not pad_silence), priority=0)
        return res[0].text

    async def main():
        server._infer_queue.start()
        return await asyncio.gather(*[ws_window(i) for i in range(12)])

    texts = asyncio.run(main())
    binding.uninstall()
    assert texts == [f"{float(i + 1):.1f}" for i in range(12)], texts   # window i's own result, none from the per-call path (-1.0)
    assert backend.calls == 0 and server.model.tower.calls == 0
    assert sum(len(b[1]) for b in record) == 12
    assert len(record) <= 3, f"12 concurrent windows should be encoded in very few batches, got {[len(b[1]) for b in record]}"
    assert binding.stats["prefetched"] == 12


def test_sse_chunks_are_prefetched_as_one_batch(monkeypatch):
    """The SSE loop (server.py:981-1008) awaits chunk after chunk; prefetch_chunks submits all of a request's chunks up front."""
    server = _make_server(monkeypatch)
    backend = FakeBackend()
    server_hook.install(server, batch_windows=False)
    server_hook.try_load_b200_encoder(server.model, factory=lambda t: backend)
    record = []
    binding = QueueBinding(server, max_wait_ms=5.0, encode_factory=_recording_factory(record))
    binding.install()
    sr, lang_code = 16000, None
    audio = (np.arange(16 * 16000) // 16000 + 1).astype(np.float32) / 1000.0     # second k holds (k + 1) / 1000
    chunk_n, overlap_n = 5 * sr, 1 * sr
    bounds, start = [], 0
    while start < len(audio):
        end = min(start + chunk_n, len(audio))
        bounds.append((start, end))
        if end >= len(audio):
            break
        start = end - overlap_n

    async def main():
        server._infer_queue.start()
        assert binding.prefetch_chunks([audio[a:b] for a, b in bounds]) == len(bounds)
        out = []
        for a, b in bounds:
            chunk = audio[a:b]
            res = await server._infer_queue.submit(lambda c=chunk: server._do_transcribe(c, sr, lang_code, False), priority=1)
            out.append(res[0].text)
        return out

    texts = asyncio.run(main())
    binding.uninstall()
    assert texts == [f"{float(a // sr + 1):.1f}" for a, _ in bounds]
    assert len(record) == 1 and len(record[0][1]) == len(bounds), "all chunks of the request in ONE encode"
    assert backend.calls == 0


def test_falls_back_when_the_window_cannot_be_prefetched(monkeypatch):
    """Other sample rates, long uploads, or a failing batch: the job runs exactly as without the binding."""
    server = _make_server(monkeypatch)
    backend = FakeBackend()
    server_hook.install(server, batch_windows=False)
    server_hook.try_load_b200_encoder(server.model, factory=lambda t: backend)

    def failing_factory(b):
        def encode(windows, _flush):
            raise RuntimeError("boom")

        return encode

    binding = QueueBinding(server, max_wait_ms=1.0, encode_factory=failing_factory)
    binding.install()

    async def one(audio, sr):
        lang_code = None
        res = await server._infer_queue.submit(lambda: server._do_transcribe(audio, sr, lang_code, False), priority=1)
        return res[0].text

    async def main():
        server._infer_queue.start()
        a = await one(np.ones(44100, np.float32), 44100)          # not 16 kHz: never prefetched
        b = await one(np.ones(40 * 16000, np.float32), 16000)     # longer than 30 s: the SDK's own chunking
        c = await one(np.ones(16000, np.float32), 16000)          # prefetched, but the batch fails
        return a, b, c

    assert asyncio.run(main()) == ("-1.0", "-1.0", "-1.0")
    binding.uninstall()
    assert backend.calls == 3 and binding.stats == {"prefetched": 1, "bypassed": 2}


def test_dual_model_routes_partials_to_the_fast_backend(monkeypatch):
    """WS partials run on the 0.6B model, flushes on the 1.7B one (server.py:1351): each window is batched on ITS backend."""
    server = _make_server(monkeypatch, fast=True)
    b_full, b_fast = FakeBackend(), FakeBackend()
    made = iter([b_full, b_fast])
    server_hook.install(server, batch_windows=False)
    assert server_hook.try_load_b200_encoder(server.model, server._fast_model, factory=lambda t: next(made)) == 2
    record = []
    binding = QueueBinding(server, max_wait_ms=20.0, encode_factory=_recording_factory(record))
    binding.install()

    async def ws_window(i, pad_silence):
        audio = np.full(16000, (i + 1) / 1000.0, dtype=np.float32)
        sr, lang_code = 16000, None
        res = await server._infer_queue.submit(lambda: server._do_transcribe(audio, sr, lang_code, False, use_fast=not pad_silence), priority=0)
        return res[0].text

    async def main():
        server._infer_queue.start()
        return await asyncio.gather(*[ws_window(i, pad_silence=(i % 4 == 0)) for i in range(8)])

    texts = asyncio.run(main())
    binding.uninstall()
    assert texts == [f"{float(i + 1):.1f}" for i in range(8)]
    by_backend = {id(b): sum(len(r[1]) for r in record if r[0] is b) for b in (b_full, b_fast)}
    assert by_backend == {id(b_full): 2, id(b_fast): 6}


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present on this box")
def test_the_queue_under_test_is_the_references_own(monkeypatch):
    server = _make_server(monkeypatch)
    assert type(server._infer_queue).__name__ == "PriorityInferQueue" and server.__name__ == "server"
    assert server.__file__.startswith(REF_SRC)
