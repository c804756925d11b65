"""Binding of the varlen window batcher beneath the reference's ``PriorityInferQueue``.

The reference runs ONE job at a time: ``PriorityInferQueue._worker`` (src/server.py:79-94) pops a job and awaits
``run_in_executor(_infer_executor, job.fn)``, where ``job.fn`` is a lambda around ``_do_transcribe(audio, sr, ...)`` built by
the three call sites -- HTTP ``transcribe`` (:625-631), the SSE chunk loop (:966-969, :985-988) and the WebSocket
``_transcribe_with_context`` (:1349-1355).  The encoder therefore only ever sees one 0.5-30 s clip.  The audio of a job is
known when the job is *submitted*, long before the single infer thread reaches it, so the binding works one level below the
queue without touching it:

* ``QueueBinding.install()`` replaces the *instance* attribute ``server._infer_queue.submit`` by a coroutine that first looks
  into the job's closure for its audio window (``audio`` / default ``c``, ``sr``, ``pad_silence`` -- the free variables of the
  three lambdas), hands the window to a ``WindowBatcher`` (batcher.py) bound to the backend of the model the job will use,
  wraps ``fn`` so that the infer thread carries the resulting future in ``server_hook._job_ctx`` while it runs the job, and
  then calls the original ``submit``.  Priorities, the heap, the worker and the executor are the reference's own.
* The batcher's collector thread encodes everything that was submitted within ``max_wait_ms`` as ONE ragged PCM batch
  (log-mel kernel + encoder on its own CUDA stream), while the infer thread is still busy with the decoder of an earlier job.
* When the infer thread reaches the job, the SDK computes its features and calls ``audio_tower.forward``; the patched forward
  (server_hook.patched_encoder) asks the job's ``Prefetched.take``, which returns the pre-computed hidden states iff they belong
  to the same backend and the frame count matches the ``feature_lens`` the SDK passes.  Anything else (introspection found no
  window, resampled / stereo / long audio, SDK-side chunking, a failed batch) falls through to the normal per-call encode:
  the result of a request never depends on the binding, only its latency does.
* ``prefetch_chunks`` lets a caller that knows all its windows up front (the SSE chunk loop: fixed 5 s chunks with 1 s overlap,
  :958-1008) submit them together; the later per-chunk ``submit`` calls find them by buffer address.

The kernels are batch-invariant (tests/test_gpu_path.py::test_encoder_batch_invariance_and_determinism), so the hidden states a
job receives are bit-identical to what its own ``encode_pcm`` call would have produced.
"""

from __future__ import annotations

import logging
import threading
from typing import Callable

import numpy as np

from . import server_hook
from .batcher import WindowBatcher

_log = logging.getLogger("qwen3_asr_b200")

TARGET_SR = 16000
MAX_PREFETCH_SAMPLES = 30 * TARGET_SR     # longer uploads go through the SDK's own splitter: not pre-encoded
MIN_CLIP_SAMPLES = 8000                   # the SDK pads shorter clips to 0.5 s before the extractor


def _window_key(a: np.ndarray):
    return (a.__array_interface__["data"][0], a.shape[0], a.dtype.str)


def job_window(fn):
    """(audio, sr, use_fast) of a job lambda of the reference's three call sites, or None.  Free variables by name:
    ``audio`` (HTTP, WS) or the default argument ``c`` (SSE), ``sr``, and ``pad_silence`` (WS: use_fast = not pad_silence)."""
    code = getattr(fn, "__code__", None)
    if code is None:
        return None
    cells = {}
    for name, cell in zip(code.co_freevars, fn.__closure__ or ()):
        try:
            cells[name] = cell.cell_contents
        except ValueError:   # empty cell
            pass
    audio = cells.get("audio")
    if audio is None and fn.__defaults__:
        audio = fn.__defaults__[0]
    sr = cells.get("sr")
    if not isinstance(audio, np.ndarray) or audio.ndim != 1 or not isinstance(sr, (int, np.integer)):
        return None
    if "pad_silence" in cells:          # _transcribe_with_context: partials on the fast model, the flush on the full one
        use_fast = not cells["pad_silence"]
    else:
        use_fast = False
    return audio, int(sr), bool(use_fast)


class Prefetched:
    """The pre-computed encoder output of one job (a future from the WindowBatcher) plus what it was computed for."""

    def __init__(self, future, backend, n_samples: int):
        self.future, self.backend, self.n_samples = future, backend, n_samples
        self.used = False

    def take(self, backend, input_features, feature_lens):
        """Called by the patched ``audio_tower.forward`` on the infer thread.  Returns an object with ``.last_hidden_state`` or
        None (caller encodes as usual).  One job = one clip = one forward: a second forward of the same job is not served."""
        if self.used or backend is not self.backend:
            return None
        self.used = True
        lens = _as_int_list(feature_lens, input_features)
        if lens is None or len(lens) != 1 or lens[0] != self.n_samples // 160:
            return None
        try:
            hidden = self.future.result(timeout=30.0)
        except Exception:    # the batch failed (or timed out): this job encodes for itself, as it would without the binding
            return None
        from .encoder import EncoderOutput

        import torch

        # computed on the batcher's stream: make the caller's stream wait for it
        ev = getattr(hidden, "_qasr_ready", None)
        if ev is not None:
            torch.cuda.current_stream(hidden.device).wait_event(ev)
        return EncoderOutput(last_hidden_state=hidden)


def _as_int_list(feature_lens, input_features):
    if feature_lens is None:
        if input_features is not None and getattr(input_features, "dim", lambda: 0)() == 2:
            return [int(input_features.shape[1])]
        return None
    try:
        return [int(v) for v in np.asarray(getattr(feature_lens, "cpu", lambda: feature_lens)()).reshape(-1)]
    except Exception:
        return None


class QueueBinding:
    def __init__(self, server_module, log=None, max_wait_ms: float = 3.0, max_audio_s: float = 960.0,
                 encode_factory: Callable | None = None):
        self.server = server_module
        self.log = log or _log
        self.max_wait_ms, self.max_audio_s = max_wait_ms, max_audio_s
        self._encode_factory = encode_factory or _default_encode
        self._batchers: dict[int, tuple] = {}     # id(backend) -> (backend, WindowBatcher)
        self._cache: dict = {}                    # window key -> Prefetched (submitted ahead of its job: SSE chunks)
        self._lock = threading.Lock()
        self._orig_submit = None
        self.stats = {"prefetched": 0, "bypassed": 0}

    # ---- wiring ------------------------------------------------------------------------------------------------
    def install(self) -> None:
        q = self.server._infer_queue
        if self._orig_submit is not None:
            return
        self._orig_submit = q.submit            # bound method of the reference's PriorityInferQueue: called, never edited

        async def submit(fn, priority: int = 1):
            return await self._orig_submit(self._wrap(fn), priority)

        q.submit = submit                        # instance attribute: shadows the class method for this one queue object

    def uninstall(self) -> None:
        if self._orig_submit is not None:
            try:
                del self.server._infer_queue.submit
            except AttributeError:
                pass
            self._orig_submit = None
        self.drain()

    def drain(self) -> None:
        """Stop the collector threads (model unload): pending windows are still encoded, then the batchers go away."""
        with self._lock:
            batchers, self._batchers = list(self._batchers.values()), {}
            self._cache.clear()
        for _, b in batchers:
            b.close()

    # ---- per job -----------------------------------------------------------------------------------------------
    def _backend_for(self, use_fast: bool):
        fast = getattr(self.server, "_fast_model", None)
        m = fast if (use_fast and fast is not None) else getattr(self.server, "model", None)
        if m is None:
            return None
        return server_hook.backend_for(m)[1]

    def _batcher(self, backend) -> WindowBatcher:
        with self._lock:
            hit = self._batchers.get(id(backend))
            if hit is None or hit[0] is not backend:
                hit = (backend, WindowBatcher(self._encode_factory(backend), max_wait_ms=self.max_wait_ms, max_audio_s=self.max_audio_s))
                self._batchers[id(backend)] = hit
            return hit[1]

    def _prefetch(self, audio: np.ndarray, sr: int, use_fast: bool):
        if sr != TARGET_SR or audio.ndim != 1 or not (MIN_CLIP_SAMPLES <= audio.shape[0] <= MAX_PREFETCH_SAMPLES):
            return None
        if getattr(self.server, "USE_SPECULATIVE", False):   # _do_transcribe_speculative runs both models (server.py:823-846)
            return None
        backend = self._backend_for(use_fast)
        if backend is None:
            return None
        fut = self._batcher(backend).submit(np.ascontiguousarray(audio, dtype=np.float32))
        return Prefetched(fut, backend, int(audio.shape[0]))

    def prefetch_chunks(self, chunks, sr: int = TARGET_SR, use_fast: bool = False) -> int:
        """Submit several windows of one request at once (e.g. all SSE chunks ``audio[start:end]`` of server.py:981-1008) so
        they are encoded as one batch; the job that later carries the same buffer (same address and length) picks its result up."""
        n = 0
        for c in chunks:
            pre = self._prefetch(c, sr, use_fast)
            if pre is not None:
                with self._lock:
                    self._cache[_window_key(c)] = pre
                n += 1
        return n

    def _wrap(self, fn):
        pre = None
        try:
            win = job_window(fn)
            if win is not None and server_hook.n_backends() > 0:
                audio, sr, use_fast = win
                with self._lock:
                    pre = self._cache.pop(_window_key(audio), None)
                if pre is None:
                    pre = self._prefetch(audio, sr, use_fast)
        except Exception as e:      # never let the binding break a request
            self.log.error(f"B200 window batching: prefetch skipped: {e}")
            pre = None
        if pre is None:
            self.stats["bypassed"] += 1
            return fn
        self.stats["prefetched"] += 1

        def job():
            server_hook._job_ctx.prefetch = pre
            try:
                return fn()
            finally:
                server_hook._job_ctx.prefetch = None

        return job


def _default_encode(backend):
    """encode(windows, flush_flags) for the WindowBatcher: float clips -> one ragged log-mel + encoder call on a stream of
    the collector thread's own, so it overlaps the infer thread's decoder work."""
    import torch

    stream = torch.cuda.Stream(device=backend.tdev)

    def encode(windows, _flush):
        with torch.cuda.stream(stream):
            hidden, toks = backend.encode_pcm(list(windows))
            ev = torch.cuda.Event()
            ev.record(stream)
        outs = _SplitWithEvent(hidden, ev)
        return outs, toks

    return encode


class _SplitWithEvent:
    """Slicing facade over the batch's hidden states: every per-window slice carries the CUDA event that marks the batch as
    computed, so the consumer's stream can wait for it (Prefetched.take)."""

    def __init__(self, hidden, event):
        self.hidden, self.event = hidden, event

    def __getitem__(self, sl):
        out = self.hidden[sl]
        out._qasr_ready = self.event
        return out
