"""Varlen batch builder for concurrent WebSocket / SSE windows.

The reference encodes one window per job: every WS trigger and every SSE chunk becomes its own ``_do_transcribe`` call
on the single inference thread (src/server.py:79-94, 1291-1375), so the encoder sees batches of one 0.5-6 s clip -- a few
hundred tokens, far below what fills a B200.  ``WindowBatcher`` is a micro-batching shim *below* ``PriorityInferQueue``
(which stays untouched): callers hand in a window and get a future; a collector thread gathers whatever arrives within
``max_wait_ms`` (or until ``max_audio_s`` of audio is pending), runs ONE ragged-batch encoder call and resolves every
future with that clip's hidden states.  The kernels are batch-invariant (tests/test_gpu_path.py::
test_encoder_batch_invariance_and_determinism), so a window's result does not depend on what it was batched with.

Pure host logic; the encode callable is injected (``B200PreFrontend.encode_windows`` for PCM bytes,
``B200AudioEncoder.encode_pcm`` for float clips), so this module is covered on CPU with a stub.
"""

from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import Callable, Sequence


class WindowBatcher:
    def __init__(self, encode: Callable[[Sequence, Sequence[bool]], tuple], max_wait_ms: float = 5.0, max_audio_s: float = 1024.0,
                 seconds_of: Callable[[object], float] | None = None, empty: Callable[[], object] | None = None):
        """encode(windows, flush_flags) -> (hidden [sum tokens, D], token_lens [n]) for the whole batch, clip-major.
        Empty windows (the reference returns '' for them before any inference, src/server.py:1331-1332) never reach ``encode``:
        their future resolves to ``hidden[:0]`` of the batch they arrived with (or ``empty()`` when the whole batch was empty),
        so one client's empty flush cannot fail the windows of other clients."""
        self._encode = encode
        self.max_wait = max_wait_ms / 1000.0
        self.max_audio_s = float(max_audio_s)
        self._empty = empty or (lambda: None)
        self._seconds_of = seconds_of or (lambda w: len(w) / (2 * 16000.0) if isinstance(w, (bytes, bytearray)) else len(w) / 16000.0)
        self._lock = threading.Condition()
        self._pending = []          # (window, flush, future, t_arrival)
        self._pending_s = 0.0
        self._stop = False
        self.batches = []           # sizes of the batches formed so far (observability / tests)
        self._thread = threading.Thread(target=self._run, name="qasr-b200-batcher", daemon=True)
        self._thread.start()

    def submit(self, window, flush: bool = False) -> Future:
        """Thread-safe.  The future resolves to this window's hidden states [tokens_i, D]."""
        fut: Future = Future()
        with self._lock:
            if self._stop:
                raise RuntimeError("WindowBatcher is closed")
            self._pending.append((window, bool(flush), fut, time.monotonic()))
            self._pending_s += self._seconds_of(window)
            self._lock.notify_all()
        return fut

    def close(self) -> None:
        with self._lock:
            self._stop = True
            self._lock.notify_all()
        self._thread.join()

    # ---- collector -------------------------------------------------------------------------------------------
    def _take(self):
        """Wait for work; return the batch to run (FIFO, capped at max_audio_s) or None at shutdown."""
        with self._lock:
            while True:
                if self._pending:
                    oldest = self._pending[0][3]
                    due = oldest + self.max_wait
                    now = time.monotonic()
                    if self._stop or self._pending_s >= self.max_audio_s or now >= due:
                        batch, total = [], 0.0
                        while self._pending:
                            s = self._seconds_of(self._pending[0][0])
                            if batch and total + s > self.max_audio_s:
                                break
                            batch.append(self._pending.pop(0))
                            total += s
                        self._pending_s = max(0.0, self._pending_s - total)
                        return batch
                    self._lock.wait(timeout=due - now)
                elif self._stop:
                    return None
                else:
                    self._lock.wait()

    def _run(self) -> None:
        while True:
            batch = self._take()
            if batch is None:
                return
            futs = [b[2] for b in batch]
            live = [f.set_running_or_notify_cancel() for f in futs]
            try:
                keep = [i for i, b in enumerate(batch) if self._seconds_of(b[0]) > 0]
                hidden, token_lens = None, []
                if keep:
                    hidden, token_lens = self._encode([batch[i][0] for i in keep], [batch[i][1] for i in keep])
                self.batches.append(len(batch))
                lens = dict(zip(keep, token_lens))
                start = 0
                for i, (f, ok) in enumerate(zip(futs, live)):
                    n = int(lens.get(i, 0))
                    if ok:
                        f.set_result(hidden[start:start + n] if hidden is not None else self._empty())
                    start += n
            except BaseException as e:  # the whole batch shares the failure, as a failing job does in the reference (server.py:93-94)
                for f, ok in zip(futs, live):
                    if ok:
                        f.set_exception(e)
