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
"""

from __future__ import annotations

import contextlib
import logging
import os

_log = logging.getLogger("qwen3_asr_b200")

# one backend per torch module (dual-model mode keeps a 1.7B and a 0.6B tower resident, server.py:411-425)
_b200_encoders: dict[int, object] = {}
_fallback_logged = False


def enabled() -> bool:
    return os.getenv("B200_ENCODER", "") not in ("", "0", "false", "False")


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
            if id(tower) in _b200_encoders:
                continue
            _b200_encoders[id(tower)] = (factory or _make_backend)(tower)
            log.info("B200 encoder backend ready")
            n += 1
        except Exception as e:  # same policy as the TRT / ONNX loaders: report and continue without it
            log.error(f"B200 encoder load failed: {e}")
    return n


def unload() -> None:
    """Free the backends (call from ``_unload_model_sync``, server.py:478-496)."""
    for enc in _b200_encoders.values():
        close = getattr(enc, "close", None)
        if close:
            close()
    _b200_encoders.clear()


def backend_for(m):
    tower = find_audio_tower(m)
    if tower is None:
        return None, None
    return tower, _b200_encoders.get(id(tower))


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


def install(server_module, log=None) -> None:
    """Zero-edit integration: wrap ``server._do_transcribe`` and ``server._load_model_sync`` in place.

    After ``install(server)``, loading the model also runs ``try_load_b200_encoder(model, _fast_model)`` and every
    ``_do_transcribe`` call runs with the selected model's audio tower patched.  With B200_ENCODER unset both
    wrappers are pass-throughs."""
    log = log or getattr(server_module, "log", None) or _log
    if getattr(server_module, "_b200_installed", False):
        return
    orig_do = server_module._do_transcribe
    orig_load = getattr(server_module, "_load_model_sync", None)

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
            try_load_b200_encoder(getattr(server_module, "model", None), getattr(server_module, "_fast_model", None), log=log)
            return r

        server_module._load_model_sync = _load_model_sync
    server_module._b200_installed = True
