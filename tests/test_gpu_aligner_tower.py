"""SURVEY 8f-4, ForcedAligner reuse: Qwen3-ForcedAligner shares the audio tower with Qwen3-ASR (vLLM's
Qwen3ASRForcedAlignerForTokenClassification subclasses Qwen3ASRForConditionalGeneration and keeps its ``audio_tower``,
vllm/model_executor/models/qwen3_asr_forced_aligner.py:30-45), and the reference loads it as a second SDK model
(src/subtitle.py:315-331).  Here the in-container tower that vLLM builds for both -- its independent re-implementation of the audio
encoder (qwen3_omni_moe_thinker.py:321-533: fused QKV, explicit cu_seqlens attention, no mask to inject) -- is stood up on the GPU in
bf16 with seeded weights and compared with the CUDA backend created from the very same weights: a second reference implementation,
written by other people, agreeing with this one."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vllm_tower(cfg, weights):
    """vLLM's audio encoder outside an engine: a 1-rank distributed environment and a default VllmConfig are enough."""
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
    from vllm.config import VllmConfig, set_current_vllm_config
    from vllm.distributed import init_distributed_environment, initialize_model_parallel
    from vllm.distributed.parallel_state import model_parallel_is_initialized
    from vllm.model_executor.models.qwen3_omni_moe_thinker import Qwen3OmniMoeAudioEncoder

    hc = Qwen3OmniMoeAudioEncoderConfig(
        num_mel_bins=128, encoder_layers=cfg.layers, encoder_attention_heads=cfg.heads, encoder_ffn_dim=cfg.ffn,
        d_model=cfg.d_model, output_dim=cfg.output_dim, n_window=50, n_window_infer=800, conv_chunksize=500,
        downsample_hidden_size=480, max_source_positions=1500, activation_function="gelu", scale_embedding=False,
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    vcfg = VllmConfig()
    with set_current_vllm_config(vcfg):
        if not model_parallel_is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29741")
            init_distributed_environment(world_size=1, rank=0, distributed_init_method="tcp://127.0.0.1:29741", local_rank=0, backend="nccl")
            initialize_model_parallel(1, 1)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            with torch.device("cuda"):
                tower = Qwen3OmniMoeAudioEncoder(hc, prefix="audio_tower").eval()
        finally:
            torch.set_default_dtype(prev)
    sd = {}
    for k, v in weights.items():
        if ".self_attn.q_proj." in k:
            base = k.replace("q_proj", "{}")
            kind = k.rsplit(".", 1)[1]
            sd[k.replace("q_proj", "qkv")] = torch.cat([weights[base.format(n)] for n in ("q_proj", "k_proj", "v_proj")], dim=0)
        elif ".self_attn.k_proj." in k or ".self_attn.v_proj." in k:
            continue
        else:
            sd[k] = v
    missing, unexpected = tower.load_state_dict({k: v.to("cuda", torch.bfloat16) for k, v in sd.items()}, strict=False)
    assert not unexpected and all("positional_embedding" in m for m in missing), (missing, unexpected)
    return tower, vcfg


@pytest.mark.parametrize("name", ["0.6B"])
def test_vllms_audio_tower_agrees_with_the_cuda_backend(name):
    """Runs in a child process: standing vLLM's modules up initialises a process-wide distributed environment that must not leak
    into the other GPU tests of this session."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    if r.returncode == 77:
        pytest.skip(tail.strip().splitlines()[-1])
    assert r.returncode == 0, tail
    print(next(line for line in r.stdout.splitlines() if line.startswith("[")))


class _Skip(Exception):
    pass


def _check(name):
    from oracle import CONFIGS, encoder_forward, make_weights, oracle_device_fp32
    from oracle.signals import speech_like
    from qwen3_asr_b200 import B200AudioEncoder

    cfg = CONFIGS[name]
    w = make_weights(cfg, seed=5)
    try:
        from vllm.config import set_current_vllm_config

        tower, vcfg = _vllm_tower(cfg, w)
    except Exception as e:  # noqa: BLE001 -- vLLM internals move between versions; the test documents what it could not build
        raise _Skip(f"vLLM's audio tower could not be stood up outside an engine here: {type(e).__name__}: {str(e)[:200]}")
    enc = B200AudioEncoder(cfg, w, max_chunks=128)
    try:
        clips = [speech_like(30 * 16000, 60), speech_like(int(6.4 * 16000), 61)]
        mel, flens = enc.logmel(clips)
        out = enc.encode(mel, flens).float()
        toks = [int(t) for t in enc.last_token_lens]
        cols = np.concatenate([[0], np.cumsum(flens)])
        mels = [mel[:, cols[i]:cols[i + 1]].to(torch.bfloat16) for i in range(len(clips))]
        with torch.inference_mode(), set_current_vllm_config(vcfg):
            ref = torch.cat([tower(m, torch.tensor([m.shape[1]], device="cuda"), torch.tensor([t], device="cuda")).float()
                             for m, t in zip(mels, toks)])
        assert ref.shape == out.shape
        w_dev = {k: v.cuda() for k, v in w.items()}
        with oracle_device_fp32():
            truth, _ = encoder_forward(w_dev, cfg, [m.float() for m in mels], device="cuda")

        def rms(a, b):
            return float(((a.double() - b.double()).pow(2).mean() / b.double().pow(2).mean()).sqrt())

        e_cuda, e_vllm, e_cross = rms(out, truth), rms(ref, truth), rms(out, ref)
        print(f"[{name}] rms-rel error vs the fp32 oracle: CUDA backend {e_cuda:.3e}, vLLM's bf16 tower {e_vllm:.3e}; CUDA vs vLLM {e_cross:.3e}")
        assert e_cuda <= 1.2 * e_vllm and e_cross <= 2.0 * e_vllm
        assert float((out - ref).abs().max() / ref.abs().max()) <= 4e-2
    finally:
        enc.close()


if __name__ == "__main__":
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        _check(sys.argv[1] if len(sys.argv) > 1 else "0.6B")
    except _Skip as e:
        print(str(e))
        sys.exit(77)
