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
