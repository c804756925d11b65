"""The drop-in boundary: the B200 slot driven through the REFERENCE's own ``_do_transcribe`` (imported from
/root/reference/src when present -- this container only) with a fake SDK model, plus the same plumbing against
a stand-alone fake server so the logic is also covered where the reference tree does not exist.  CPU only: the
backend object is a recording fake; the real CUDA backend behind the same hook is covered by
tests/test_gpu_path.py::test_hook_with_transformers_audio_tower."""

import os
import sys
import types

import numpy as np
import pytest
import torch

from qwen3_asr_b200 import server_hook

REF_SRC = "/root/reference/src"


class FakeTower(torch.nn.Module):
    """Shaped like Qwen3OmniMoeAudioEncoder: forward(input_features, feature_lens=...) -> obj.last_hidden_state"""

    def __init__(self):
        super().__init__()
        self.calls = 0

    def forward(self, input_features, feature_lens=None, aftercnn_lens=None):
        self.calls += 1
        return types.SimpleNamespace(last_hidden_state=torch.zeros(3, 4))


class FakeSDKModel:
    """qwen_asr.Qwen3ASRModel layout: .model.thinker.audio_tower, .transcribe((audio, sr), ...)"""

    def __init__(self):
        self.tower = FakeTower()
        self.model = types.SimpleNamespace(thinker=types.SimpleNamespace(audio_tower=self.tower))
        self.seen = []

    def transcribe(self, audio_sr, language=None, return_time_stamps=False):
        out = self.model.thinker.audio_tower.forward(torch.ones(128, 50), feature_lens=torch.tensor([50]))
        self.seen.append(out.last_hidden_state.clone())
        return [types.SimpleNamespace(text="hello", language="en")]


class FakeBackend:
    def __init__(self, fail=False):
        self.fail = fail
        self.calls = 0
        self.closed = False

    def forward(self, input_features, feature_lens=None):
        self.calls += 1
        if self.fail:
            raise RuntimeError("boom")
        return types.SimpleNamespace(last_hidden_state=torch.full((3, 4), 7.0))

    def close(self):
        self.closed = True


@pytest.fixture(autouse=True)
def _clean(monkeypatch):
    server_hook.unload()
    server_hook._fallback_logged = False
    monkeypatch.delenv("B200_ENCODER", raising=False)
    yield
    server_hook.unload()


def test_unset_env_means_unchanged_behaviour():
    m = FakeSDKModel()
    assert server_hook.try_load_b200_encoder(m, factory=lambda t: FakeBackend()) == 0
    with server_hook.patched_encoder(m) as active:
        assert active is False
        m.transcribe((np.zeros(10), 16000))
    assert m.tower.calls == 1


def test_loader_failure_is_logged_not_raised(monkeypatch):
    monkeypatch.setenv("B200_ENCODER", "1")
    errors = []
    log = types.SimpleNamespace(error=errors.append, info=lambda *_: None)

    def bad_factory(tower):
        raise RuntimeError("no GPU")

    assert server_hook.try_load_b200_encoder(FakeSDKModel(), log=log, factory=bad_factory) == 0
    assert errors and "no GPU" in errors[0]


def test_patch_routes_restores_and_falls_back(monkeypatch):
    monkeypatch.setenv("B200_ENCODER", "1")
    m = FakeSDKModel()
    backend = FakeBackend()
    assert server_hook.try_load_b200_encoder(m, factory=lambda t: backend) == 1
    orig = m.tower.forward
    res = server_hook.run_transcribe(m, lambda: m.transcribe((np.zeros(10), 16000)))
    assert res[0].text == "hello"
    assert backend.calls == 1 and m.tower.calls == 0
    assert float(m.seen[-1][0, 0]) == 7.0
    assert m.tower.forward == orig, "forward must be restored in finally"
    # backend raises -> silent per-call fallback to the original forward (TRT-slot convention), logged once
    backend.fail = True
    errors = []
    log = types.SimpleNamespace(error=errors.append, info=lambda *_: None)
    server_hook.run_transcribe(m, lambda: m.transcribe((np.zeros(10), 16000)), log=log)
    server_hook.run_transcribe(m, lambda: m.transcribe((np.zeros(10), 16000)), log=log)
    assert m.tower.calls == 2 and float(m.seen[-1][0, 0]) == 0.0
    assert len(errors) == 1
    # restore even when transcribe itself raises
    with pytest.raises(ZeroDivisionError):
        server_hook.run_transcribe(m, lambda: 1 / 0)
    assert m.tower.forward == orig
    server_hook.unload()
    assert backend.closed


def test_dual_model_picks_backend_by_module_identity(monkeypatch):
    monkeypatch.setenv("B200_ENCODER", "1")
    full, fast = FakeSDKModel(), FakeSDKModel()
    b_full, b_fast = FakeBackend(), FakeBackend()
    made = iter([b_full, b_fast])
    assert server_hook.try_load_b200_encoder(full, fast, factory=lambda t: next(made)) == 2
    server_hook.run_transcribe(fast, lambda: fast.transcribe((np.zeros(10), 16000)))
    assert (b_full.calls, b_fast.calls) == (0, 1)


def test_legacy_encoder_attribute_returns_tuple(monkeypatch):
    """The reference's branches patch ``m.encoder.forward`` and expect ``(out,)`` (server.py:876-882)."""
    monkeypatch.setenv("B200_ENCODER", "1")
    enc_mod = FakeTower()
    m = types.SimpleNamespace(encoder=enc_mod)
    server_hook.try_load_b200_encoder(m, factory=lambda t: FakeBackend())
    with server_hook.patched_encoder(m):
        out = m.encoder.forward(torch.ones(1, 128, 30))
        assert isinstance(out, tuple) and float(out[0][0, 0]) == 7.0
        assert m.encoder.forward(None) is not None  # input None -> original forward


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present on this box")
def test_through_the_references_own_do_transcribe(monkeypatch):
    """Import the reference's server.py unmodified, install the slot, and drive its _do_transcribe."""
    monkeypatch.setenv("B200_ENCODER", "1")
    monkeypatch.setenv("MODEL_ID", "Qwen/Qwen3-ASR-1.7B")
    monkeypatch.syspath_prepend(REF_SRC)
    sys.modules.pop("server", None)
    try:
        import server  # noqa: the reference module
    except Exception as e:  # optional dependency of the reference missing here
        pytest.skip(f"reference server.py not importable: {e}")
    m = FakeSDKModel()
    backend = FakeBackend()
    monkeypatch.setattr(server, "model", m, raising=False)
    monkeypatch.setattr(server, "_fast_model", None, raising=False)
    monkeypatch.setattr(server, "_cuda_stream", None, raising=False)
    monkeypatch.setattr(server, "_PINNED_AUDIO_BUFFER", None, raising=False)
    server_hook.install(server)
    server_hook.try_load_b200_encoder(server.model, factory=lambda t: backend)
    audio = np.zeros(16000, dtype=np.float32)
    res = server._do_transcribe(audio, 16000, None, False)
    assert res[0].text == "hello"
    assert backend.calls == 1 and m.tower.calls == 0, "the reference's _do_transcribe must reach the B200 backend"
    assert float(m.seen[-1][0, 0]) == 7.0
    # and with the slot disabled the same call goes to the torch encoder
    server_hook.unload()
    server._do_transcribe(audio, 16000, None, False)
    assert m.tower.calls == 1
    sys.modules.pop("server", None)


def test_forced_aligner_reuses_the_same_slot(monkeypatch):
    """the ForcedAligner (src/subtitle.py:315-331) is the same audio-tower family behind the same SDK layout
    (.model.thinker.audio_tower), so the same loader / patch serves /v1/audio/subtitles: a third model gets its own backend and
    only its own tower is patched while it runs."""
    monkeypatch.setenv("B200_ENCODER", "1")
    asr, aligner = FakeSDKModel(), FakeSDKModel()
    aligner.align = lambda audio_sr, text, language=None: aligner.transcribe(audio_sr)     # Qwen3ForcedAligner.align runs the tower too
    b_asr, b_al = FakeBackend(), FakeBackend()
    made = iter([b_asr, b_al])
    assert server_hook.try_load_b200_encoder(asr, factory=lambda t: next(made)) == 1
    assert server_hook.try_load_b200_encoder(aligner, factory=lambda t: next(made)) == 1   # lazy load_aligner() calls it again
    orig = aligner.tower.forward
    server_hook.run_transcribe(aligner, lambda: aligner.align((np.zeros(10), 16000), "hello"))
    assert (b_asr.calls, b_al.calls) == (0, 1)
    assert aligner.tower.forward == orig and aligner.tower.calls == 0                      # restored; the torch forward never ran
    assert float(aligner.seen[0][0, 0]) == 7.0


def _fake_server(m, fast=None):
    """The names install() wraps, as a stand-alone module object (the reference's server.py is used where it is importable)."""
    server = types.SimpleNamespace(model=None, _fast_model=None, loads=0, unloads=0)

    def _load_model_sync():
        server.model, server._fast_model = m(), (fast() if fast else None)
        server.loads += 1

    def _unload_model_sync():            # src/server.py:478-496
        server.model, server._fast_model = None, None
        server.unloads += 1

    def _do_transcribe(audio, sr, lang_code, return_timestamps, use_fast=False):
        mm = server._fast_model if (use_fast and server._fast_model is not None) else server.model
        return mm.transcribe((audio, sr), language=lang_code, return_time_stamps=return_timestamps)

    server._load_model_sync, server._unload_model_sync, server._do_transcribe = _load_model_sync, _unload_model_sync, _do_transcribe
    return server


def test_idle_unload_frees_the_backend_and_reload_gets_a_new_one(monkeypatch):
    """install() must wrap _unload_model_sync (idle timeout, server.py:478-527) -- otherwise ~6 GB of weights + workspace
    leak per idle cycle -- and a reloaded model (a NEW tower object) must get a NEW backend, never a stale one."""
    monkeypatch.setenv("B200_ENCODER", "1")
    made = []

    def factory(tower):
        made.append(FakeBackend())
        return made[-1]

    monkeypatch.setattr(server_hook, "_make_backend", factory)
    server = _fake_server(FakeSDKModel)
    server_hook.install(server)
    server._load_model_sync()
    assert len(made) == 1 and server_hook.n_backends() == 1
    server._do_transcribe(np.zeros(16000, np.float32), 16000, None, False)
    assert made[0].calls == 1
    server._unload_model_sync()
    assert server.unloads == 1 and made[0].closed and server_hook.n_backends() == 0
    server._load_model_sync()
    assert len(made) == 2 and not made[1].closed
    server._do_transcribe(np.zeros(16000, np.float32), 16000, None, False)
    assert (made[0].calls, made[1].calls) == (1, 1)


def test_a_garbage_collected_tower_closes_its_backend_and_never_matches_again(monkeypatch):
    import gc

    monkeypatch.setenv("B200_ENCODER", "1")
    m = FakeSDKModel()
    backend = FakeBackend()
    server_hook.try_load_b200_encoder(m, factory=lambda t: backend)
    assert server_hook.backend_for(m)[1] is backend
    del m
    gc.collect()
    assert backend.closed and server_hook.n_backends() == 0
    # a new model (possibly at the recycled address) resolves to nothing until it is loaded itself
    m2 = FakeSDKModel()
    assert server_hook.backend_for(m2)[1] is None


def test_unload_of_one_model_leaves_the_others(monkeypatch):
    """subtitle.unload_aligner (subtitle.py:334-341) frees only the aligner's backend."""
    monkeypatch.setenv("B200_ENCODER", "1")
    asr, aligner = FakeSDKModel(), FakeSDKModel()
    b_asr, b_al = FakeBackend(), FakeBackend()
    made = iter([b_asr, b_al])
    server_hook.try_load_b200_encoder(asr, aligner, factory=lambda t: next(made))
    sub = types.SimpleNamespace(_aligner=aligner, unload_aligner=lambda: setattr(sub, "_aligner", None))
    server = _fake_server(FakeSDKModel)
    server_hook.install(server, subtitle_module=sub)
    sub.unload_aligner()
    assert sub._aligner is None and b_al.closed and not b_asr.closed
    assert server_hook.backend_for(asr)[1] is b_asr


class FakeExtractor:
    """the attributes and call signature of WhisperFeatureExtractor the processor relies on"""

    sampling_rate, hop_length, n_fft, feature_size = 16000, 160, 400, 128

    def __init__(self):
        self.calls = 0

    def __call__(self, raw_speech, **kw):
        self.calls += 1
        raise AssertionError("the CPU extractor must not run once the CUDA frontend is installed")


class FakeMelBackend(FakeBackend):
    """backend with the log-mel entry: logmel(clips) -> (mel [128, sum T], feature_lens)"""

    def __init__(self):
        super().__init__()
        self.mel_calls = 0

    def logmel(self, clips):
        self.mel_calls += 1
        flens = np.array([len(c) // 160 for c in clips], dtype=np.int64)
        return torch.full((128, int(flens.sum())), 0.25), flens


class FakeSDKModelWithProcessor(FakeSDKModel):
    """... whose transcribe runs processor.feature_extractor first, the way the SDK processor does
    (vllm transformers_utils/processors/qwen3_asr.py:114-130), then moves the BatchFeature like ``inputs.to(device)``"""

    def __init__(self):
        super().__init__()
        self.processor = types.SimpleNamespace(feature_extractor=FakeExtractor())

    def transcribe(self, audio_sr, language=None, return_time_stamps=False):
        audio, sr = audio_sr
        feats = self.processor.feature_extractor([audio], sampling_rate=16000, padding=True, truncation=False,
                                                 return_attention_mask=True, return_tensors="pt")
        feats["feature_attention_mask"] = feats.pop("attention_mask")
        feats = feats.to("cpu")
        lens = feats["feature_attention_mask"].sum(-1)
        packed = feats["input_features"].permute(0, 2, 1)[feats["feature_attention_mask"].bool()].permute(1, 0)
        out = self.model.thinker.audio_tower.forward(packed, feature_lens=lens)
        self.seen.append((packed.clone(), out.last_hidden_state.clone()))
        return [types.SimpleNamespace(text="hello", language="en")]


def test_install_frontend_puts_the_mel_kernel_on_the_request_path(monkeypatch):
    """install(server, frontend=True) swaps m.processor.feature_extractor for B200FeatureExtractor, which returns a
    BatchFeature (pop / item assignment / .to()) shaped like the extractor's; unload restores nothing it should not."""
    monkeypatch.setenv("B200_ENCODER", "1")
    backend = FakeMelBackend()
    monkeypatch.setattr(server_hook, "_make_backend", lambda tower: backend)
    server = _fake_server(FakeSDKModelWithProcessor)
    server_hook.install(server, frontend=True)
    server._load_model_sync()
    from qwen3_asr_b200.frontend import B200FeatureExtractor

    fe = server.model.processor.feature_extractor
    assert isinstance(fe, B200FeatureExtractor) and isinstance(fe.original, FakeExtractor)
    res = server._do_transcribe(np.zeros(16000 * 2 + 77, np.float32), 16000, None, False)
    assert res[0].text == "hello"
    packed, hidden = server.model.seen[-1]
    assert backend.mel_calls == 1 and backend.calls == 1 and fe.original.calls == 0
    assert tuple(packed.shape) == (128, 200) and float(packed[0, 0]) == 0.25 and float(hidden[0, 0]) == 7.0
    # idempotent, reversible
    assert server_hook.install_frontend(server.model) is True
    server_hook.uninstall_frontend(server.model)
    assert isinstance(server.model.processor.feature_extractor, FakeExtractor)


def test_feature_extractor_rejects_what_the_kernel_does_not_implement():
    from qwen3_asr_b200.frontend import B200FeatureExtractor

    bad = FakeExtractor()
    bad.feature_size = 80     # the reference's own exporters trace with 80 bins: not this model
    with pytest.raises(ValueError):
        B200FeatureExtractor(FakeMelBackend(), original=bad)
    fe = B200FeatureExtractor(FakeMelBackend())
    with pytest.raises(ValueError):
        fe([np.zeros(1600, np.float32)], sampling_rate=8000)
    with pytest.raises(ValueError):
        fe([np.zeros(100, np.float32)], sampling_rate=16000)
    out = fe([np.zeros(1600, np.float32)], sampling_rate=16000)          # default return_tensors=None -> numpy, like the original
    assert isinstance(out["input_features"], np.ndarray) and out["input_features"].shape == (1, 128, 10)
