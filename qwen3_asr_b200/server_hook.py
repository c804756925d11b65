"""The third encoder-backend slot of jaaacki/qwen3-asr's ``src/server.py``.

The reference has two optional encoder backends, both wired the same way:

* a loader run once at model load that reads an env var and fills a module global
  (``_try_load_trt_encoder`` server.py:237-251 / ``_try_load_onnx_encoder`` :461-475, called at :407-409);
  unset variable => nothing happens; any failure => ``log.error`` and the server carries on without it;
* a branch in ``_do_transcribe`` (:873-893 / :895-914) that saves the encoder's ``forward``, installs a
  replacement that feeds ``input_features`` to the backend, runs ``m.transcribe`` under the server's CUDA
  stream, and restores ``forward`` in ``finally``.  The TRT branch falls back to the original forward when
  the backend raises (:880-884).

This module is that pair for the B200 backend.  Differences forced by the real SDK object:
the module that owns the encoder is ``m.model.thinker.audio_tower`` (the reference's ``hasattr(m, 'encoder')``
gate is false for the SDK wrapper), and its forward takes ``(input_features, feature_lens=...)`` and returns
an object with ``.last_hidden_state``.  Both shapes are supported.

Env vars (same convention as TRT_ENCODER_PATH / ONNX_ENCODER_PATH: unset => behaviour unchanged):
    B200_ENCODER=1              enable the slot
    B200_ENCODER_LIB=/path.so   optional: library path (default: the in-tree libqasr_b200.so)
    B200_FRONTEND=1             also replace the SDK processor's CPU WhisperFeatureExtractor by the CUDA log-mel kernel
    B200_BATCH_WINDOWS=1        also encode concurrently queued windows as one ragged batch (queue_binding.py)

Lifetime: the reference unloads the model after IDLE_TIMEOUT (``_unload_model_sync`` server.py:478-496, ``unload_aligner``
subtitle.py:334-341) and loads a NEW module object on the next request.  ``install()`` wraps both unload functions so the
backends (weights + workspace, ~6 GB for 1.7B) are freed with the model, and the registry holds the tower by weak reference:
a tower that was garbage-collected closes its backend, and a recycled ``id()`` can never resolve to a stale backend.
"""

from __future__ import annotations

import contextlib
import logging
import os
import threading
import weakref

_log = logging.getLogger("qwen3_asr_b200")


class _Entry:
    """One backend bound to one live torch module (held weakly; modules that cannot be weakly referenced are held strongly)."""

    def __init__(self, tower, backend):
        self.backend = backend
        try:
            self._ref = weakref.ref(tower)
            self._strong = None
            # the tower was freed without an explicit unload(): release the backend's device memory with it
            self._fin = weakref.finalize(tower, _close_backend, backend)
        except TypeError:
            self._ref, self._strong, self._fin = None, tower, None

    def tower(self):
        return self._strong if self._ref is None else self._ref()

    def close(self):
        if self._fin is not None:
            self._fin.detach()
        _close_backend(self.backend)


def _close_backend(backend) -> None:
    close = getattr(backend, "close", None)
    if close:
        close()


# one backend per torch module (dual-model mode keeps a 1.7B and a 0.6B tower resident, server.py:411-425)
_b200_encoders: list[_Entry] = []
_fallback_logged = False
# per-thread context of the job the infer thread is running: a pre-computed encoder result for it (queue_binding.py)
_job_ctx = threading.local()


def _flag(name: str) -> bool:
    return os.getenv(name, "") not in ("", "0", "false", "False")


def enabled() -> bool:
    return _flag("B200_ENCODER")


def _lookup(tower):
    """The live entry bound to exactly this module object (dead entries are dropped on the way)."""
    found = None
    for e in list(_b200_encoders):
        t = e.tower()
        if t is None:
            _b200_encoders.remove(e)   # its finalizer has already closed the backend
        elif t is tower:
            found = e
    return found


def find_audio_tower(m):
    """The module whose ``forward`` is the encoder: SDK wrapper -> HF model -> thinker -> audio_tower,
    or the ``.encoder`` attribute the reference's branches look for."""
    for path in (("model", "thinker", "audio_tower"), ("thinker", "audio_tower"), ("audio_tower",), ("encoder",)):
        obj = m
        for name in path:
            obj = getattr(obj, name, None)
            if obj is None:
                break
        if obj is not None and hasattr(obj, "forward"):
            return obj
    return None


def _make_backend(tower):
    from .encoder import B200AudioEncoder

    return B200AudioEncoder.from_module(tower)


def try_load_b200_encoder(*models, log=None, factory=None) -> int:
    """Loader, to be called next to ``_try_load_trt_encoder()`` (server.py:409) with the loaded model(s).
    Returns the number of backends created.  Never raises."""
    log = log or _log
    if not enabled():
        return 0
    lib = os.getenv("B200_ENCODER_LIB", "")
    if lib:
        os.environ["QASR_B200_LIB"] = lib
    n = 0
    for m in models:
        if m is None:
            continue
        try:
            tower = find_audio_tower(m)
            if tower is None:
                log.error("B200 encoder: no audio tower found on the model object")
                continue
            if _lookup(tower) is not None:
                continue
            _b200_encoders.append(_Entry(tower, (factory or _make_backend)(tower)))
            log.info("B200 encoder backend ready")
            n += 1
        except Exception as e:  # same policy as the TRT / ONNX loaders: report and continue without it
            log.error(f"B200 encoder load failed: {e}")
    return n


def unload(*models) -> None:
    """Free the backends: all of them (``_unload_model_sync``, server.py:478-496), or only those of the given model objects
    (``unload_aligner``, subtitle.py:334-341)."""
    if not models:
        doomed = list(_b200_encoders)
    else:
        towers = [find_audio_tower(m) for m in models if m is not None]
        doomed = [e for e in _b200_encoders if any(e.tower() is t for t in towers if t is not None)]
    for e in doomed:
        e.close()
        _b200_encoders.remove(e)


def backend_for(m):
    tower = find_audio_tower(m)
    if tower is None:
        return None, None
    e = _lookup(tower)
    return tower, (e.backend if e is not None else None)


def n_backends() -> int:
    return sum(1 for e in _b200_encoders if e.tower() is not None)


@contextlib.contextmanager
def patched_encoder(m, log=None):
    """Install the backend as the tower's ``forward`` for the duration of one ``m.transcribe`` call and
    restore the original in ``finally`` -- the structure of server.py:874-893."""
    global _fallback_logged
    log = log or _log
    tower, enc = backend_for(m)
    if enc is None:
        yield False
        return
    _orig_fwd = tower.forward
    legacy = tower is getattr(m, "encoder", None)  # the reference's own calling convention: returns (out,)

    def _b200_encoder_fwd(*args, **kwargs):
        global _fallback_logged
        inp = args[0] if args else kwargs.get("input_features")
        if inp is None:
            return _orig_fwd(*args, **kwargs)
        try:
            feature_lens = kwargs.get("feature_lens", args[1] if len(args) > 1 else None)
            pre = getattr(_job_ctx, "prefetch", None)
            out = pre.take(enc, inp, feature_lens) if pre is not None else None   # encoded ahead, batched with its queue neighbours
            if out is None:
                out = enc.forward(inp, feature_lens=feature_lens)
            return (out.last_hidden_state,) if legacy else out
        except Exception as e:  # TRT-slot convention: fall back to the original forward for this call
            if not _fallback_logged:
                log.error(f"B200 encoder failed, falling back to the torch encoder: {e}")
                _fallback_logged = True
            return _orig_fwd(*args, **kwargs)

    tower.forward = _b200_encoder_fwd
    try:
        yield True
    finally:
        tower.forward = _orig_fwd


def run_transcribe(m, run, cuda_stream=None, log=None):
    """Body of the new first branch of ``_do_transcribe``: patch, run under the server's stream, sync, restore."""
    with patched_encoder(m, log=log):
        if cuda_stream is not None:
            import torch

            with torch.cuda.stream(cuda_stream):
                results = run()
            cuda_stream.synchronize()
        else:
            results = run()
    return results


def install_frontend(m, log=None) -> bool:
    """Replace ``m.processor.feature_extractor`` (the SDK processor's CPU WhisperFeatureExtractor, called at
    vllm/transformers_utils/processors/qwen3_asr.py:114-130) by the CUDA log-mel kernel of this model's backend, so the mel
    kernel is on the request path (SURVEY.md 8b "optional fused frontend entry").  Idempotent; False if there is nothing to patch."""
    from .frontend import B200FeatureExtractor

    _, enc = backend_for(m)
    proc = getattr(m, "processor", None)
    if enc is None or proc is None or not hasattr(proc, "feature_extractor"):
        return False
    fe = proc.feature_extractor
    if isinstance(fe, B200FeatureExtractor):
        return True
    proc.feature_extractor = B200FeatureExtractor(enc, original=fe)
    (log or _log).info("B200 log-mel frontend installed")
    return True


def uninstall_frontend(m) -> None:
    from .frontend import B200FeatureExtractor

    proc = getattr(m, "processor", None)
    fe = getattr(proc, "feature_extractor", None)
    if isinstance(fe, B200FeatureExtractor) and fe.original is not None:
        proc.feature_extractor = fe.original


def install(server_module, log=None, frontend: bool | None = None, batch_windows: bool | None = None, subtitle_module=None) -> None:
    """Zero-edit integration: wrap the reference's module-level functions in place.

    After ``install(server)``:
      * ``_load_model_sync`` also runs ``try_load_b200_encoder(model, _fast_model)`` (and, with ``frontend`` / B200_FRONTEND=1,
        swaps the processors' feature extractor for the CUDA log-mel);
      * ``_unload_model_sync`` (and ``subtitle.unload_aligner``) also free the backends, so the idle unload releases the VRAM;
      * every ``_do_transcribe`` runs with the selected model's audio tower patched;
      * with ``batch_windows`` / B200_BATCH_WINDOWS=1, ``_infer_queue.submit`` encodes the window of every queued job ahead of
        time, batched with whatever else is queued (queue_binding.QueueBinding) -- PriorityInferQueue itself is untouched.
    With B200_ENCODER unset every wrapper is a pass-through."""
    log = log or getattr(server_module, "log", None) or _log
    if getattr(server_module, "_b200_installed", False):
        return
    frontend = _flag("B200_FRONTEND") if frontend is None else frontend
    batch_windows = _flag("B200_BATCH_WINDOWS") if batch_windows is None else batch_windows
    orig_do = server_module._do_transcribe
    orig_load = getattr(server_module, "_load_model_sync", None)
    orig_unload = getattr(server_module, "_unload_model_sync", None)

    def _do_transcribe(audio, sr, lang_code, return_timestamps, use_fast=False):
        fast = getattr(server_module, "_fast_model", None)
        m = fast if (use_fast and fast is not None) else getattr(server_module, "model", None)
        if m is None or not _b200_encoders:
            return orig_do(audio, sr, lang_code, return_timestamps, use_fast)
        with patched_encoder(m, log=log):
            return orig_do(audio, sr, lang_code, return_timestamps, use_fast)

    server_module._do_transcribe = _do_transcribe
    if orig_load is not None:
        def _load_model_sync(*a, **kw):
            r = orig_load(*a, **kw)
            models = (getattr(server_module, "model", None), getattr(server_module, "_fast_model", None))
            try_load_b200_encoder(*models, log=log)
            if frontend:
                for m in models:
                    if m is not None:
                        try:
                            install_frontend(m, log=log)
                        except Exception as e:   # loader policy: report and carry on with the CPU extractor
                            log.error(f"B200 log-mel frontend not installed: {e}")
            return r

        server_module._load_model_sync = _load_model_sync
    if orig_unload is not None:
        def _unload_model_sync(*a, **kw):
            binding = getattr(server_module, "_b200_queue_binding", None)
            if binding is not None:
                binding.drain()
            try:
                return orig_unload(*a, **kw)
            finally:
                unload()   # model, fast model and aligner are gone (server.py:478-496): free their backends with them

        server_module._unload_model_sync = _unload_model_sync
    if subtitle_module is None:
        import sys

        subtitle_module = sys.modules.get("subtitle")
    if subtitle_module is not None and hasattr(subtitle_module, "unload_aligner"):
        orig_unload_aligner = subtitle_module.unload_aligner

        def unload_aligner(*a, **kw):
            al = getattr(subtitle_module, "_aligner", None)
            if al is not None:
                unload(al)
            return orig_unload_aligner(*a, **kw)

        subtitle_module.unload_aligner = unload_aligner
    if batch_windows and hasattr(server_module, "_infer_queue"):
        from .queue_binding import QueueBinding

        server_module._b200_queue_binding = QueueBinding(server_module, log=log)
        server_module._b200_queue_binding.install()
    server_module._b200_installed = True
