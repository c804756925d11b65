"""Audio-encoder oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

A functional float32 CPU restatement of ``Qwen3OmniMoeAudioEncoder.forward``
(transformers 5.5.0, models/qwen3_omni_moe/modeling_qwen3_omni_moe.py; line
numbers below refer to that file) with per-clip standalone semantics
(SURVEY.md appendix A.3, B.1-B.6):

* ``token_len``       <- _get_feat_extract_output_lengths, :145-153
* ``chunk_plan``      <- chunk split, :711-726 (pad_to = 100 if T >= 100 else T,
                         appendix B.3 -- what a single-clip request sees)
* ``sinusoid_table``  <- SinusoidsPositionEmbedding, :88-106
* ``window_lens``     <- cu_seqlens construction, :745-752
* ``encoder_forward`` <- forward, :698-766; layer :568-616; attention :496-565
                         with the block-diagonal window mask of
                         _prepare_attention_mask (:676-693) applied, which the
                         reference deployment gets from flash-attn varlen.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class EncoderConfig:
    d_model: int
    layers: int
    heads: int
    ffn: int
    output_dim: int
    n_window: int = 50
    n_window_infer: int = 800
    downsample_hidden: int = 480
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    name: str = ""

    @property
    def chunk_frames(self) -> int:
        return self.n_window * 2

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads

    @property
    def freq_after_cnn(self) -> int:
        return (((self.num_mel_bins + 1) // 2 + 1) // 2 + 1) // 2

    def to_dict(self):
        return asdict(self)


# SURVEY.md appendix A.1
CONFIGS = {
    "1.7B": EncoderConfig(1024, 24, 16, 4096, 2048, name="1.7B"),
    "0.6B": EncoderConfig(896, 18, 14, 3584, 1024, name="0.6B"),
    # small shapes for fast CPU tests; same structure (head_dim 64, 480 conv channels)
    "tiny": EncoderConfig(128, 2, 2, 256, 192, name="tiny"),
}


def _conv_len(n: int) -> int:
    return (n - 1) // 2 + 1


def token_len(t: int) -> int:
    """:145-153.  Python floor semantics match torch's for the r = 0 case (-> 0)."""
    t = int(t)
    r = t % 100
    feat = (r - 1) // 2 + 1
    return ((feat - 1) // 2 + 1 - 1) // 2 + 1 + (t // 100) * 13


def chunk_plan(t: int, chunk_frames: int = 100):
    """Chunks of one clip as (start_frame, valid_frames, padded_frames)."""
    t = int(t)
    n_full, r = divmod(t, chunk_frames)
    pad_to = chunk_frames if t >= chunk_frames else t
    plan = [(i * chunk_frames, chunk_frames, chunk_frames) for i in range(n_full)]
    if r:
        plan.append((n_full * chunk_frames, r, pad_to))
    return plan


def window_lens(n_tokens: int, cfg: EncoderConfig, t: int | None = None):
    """:745-752 for one clip: windows of 13 * (n_window_infer // 100) tokens + remainder."""
    tok_per_chunk = token_len(cfg.chunk_frames)
    if t is not None and t < cfg.chunk_frames:
        tok_per_chunk = token_len(t)  # padded_mask_after_cnn.shape[-1] for a lone short clip
    w = tok_per_chunk * (cfg.n_window_infer // cfg.chunk_frames)
    out = [w] * (n_tokens // w)
    if n_tokens % w:
        out.append(n_tokens % w)
    return out


def sinusoid_table(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2).float())
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def weight_shapes(cfg: EncoderConfig):
    """State-dict names and shapes."""
    c, d, f = cfg.downsample_hidden, cfg.d_model, cfg.ffn
    shapes = {
        "conv2d1.weight": (c, 1, 3, 3), "conv2d1.bias": (c,),
        "conv2d2.weight": (c, c, 3, 3), "conv2d2.bias": (c,),
        "conv2d3.weight": (c, c, 3, 3), "conv2d3.bias": (c,),
        "conv_out.weight": (d, c * cfg.freq_after_cnn),
    }
    for i in range(cfg.layers):
        p = f"layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            shapes[p + f"self_attn.{nm}.weight"] = (d, d)
            shapes[p + f"self_attn.{nm}.bias"] = (d,)
        shapes[p + "self_attn_layer_norm.weight"] = (d,)
        shapes[p + "self_attn_layer_norm.bias"] = (d,)
        shapes[p + "fc1.weight"] = (f, d)
        shapes[p + "fc1.bias"] = (f,)
        shapes[p + "fc2.weight"] = (d, f)
        shapes[p + "fc2.bias"] = (d,)
        shapes[p + "final_layer_norm.weight"] = (d,)
        shapes[p + "final_layer_norm.bias"] = (d,)
    shapes["ln_post.weight"] = (d,)
    shapes["ln_post.bias"] = (d,)
    shapes["proj1.weight"] = (d, d)
    shapes["proj1.bias"] = (d,)
    shapes["proj2.weight"] = (cfg.output_dim, d)
    shapes["proj2.bias"] = (cfg.output_dim,)
    return shapes


def make_weights(cfg: EncoderConfig, seed: int = 0, sharpen: float = 3.0, bf16_exact: bool = True):
    """Seeded random weights at the given dims, as float32 tensors.

    Scaled so every activation is O(1) and attention is peaky (``sharpen``
    multiplies q/k weights): random-init at std 0.02 makes attention near-uniform
    and hides mask / softmax errors ("Weights").  With
    ``bf16_exact`` every value is representable in bf16, so the CUDA path's bf16
    copies hold exactly the numbers the float32 oracle uses.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in weight_shapes(cfg).items():
        if name.endswith("layer_norm.weight") or name == "ln_post.weight":
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            w = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            if name == "conv2d1.weight":
                w = w * 2.0
            if ".q_proj.weight" in name or ".k_proj.weight" in name:
                w = w * sharpen
        if bf16_exact:
            w = w.to(torch.bfloat16).to(torch.float32)
        out[name] = w.contiguous()
    return out


def _rnd(x: torch.Tensor, emulate_bf16: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if emulate_bf16 else x


E4M3_MAX = 448.0
AMAX_EPS = 1e-12


def _e4m3(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """x / scale rounded to float8_e4m3fn (round-to-nearest-even, saturating), returned as float32."""
    return (x / scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).to(torch.float32)


def linear_fp8(x: torch.Tensor, weight: torch.Tensor, bias, mode: str) -> torch.Tensor:
    """nn.Linear under torchao's Float8DynamicActivationFloat8WeightConfig (reference src/server.py:362-371), emulated:
    e4m3 weight and dynamically quantised e4m3 activation, fp32 accumulate, y = (xq @ wq^T) * sx * sw + bias.
    ``mode``: "per_tensor" (one scale per tensor) or "per_row" (per token / per output channel).  torchao itself is
    not installable offline: this restates its documented recipe (scale = max(amax, 1e-12) / 448) -- PARITY UNPINNED."""
    x = x.float()
    w = weight.float()
    if mode == "per_row":
        sx = x.abs().amax(dim=1, keepdim=True).clamp(min=AMAX_EPS) / E4M3_MAX
        sw = w.abs().amax(dim=1, keepdim=True).clamp(min=AMAX_EPS) / E4M3_MAX
    elif mode == "per_tensor":
        sx = (x.abs().amax().clamp(min=AMAX_EPS) / E4M3_MAX).reshape(1, 1)
        sw = (w.abs().amax().clamp(min=AMAX_EPS) / E4M3_MAX).reshape(1, 1)
    else:
        raise ValueError(mode)
    y = (_e4m3(x, sx) @ _e4m3(w, sw).T) * sx * sw.T
    return y if bias is None else y + bias


def _linear(x, weight, bias, fp8):
    return F.linear(x, weight, bias) if fp8 is None else linear_fp8(x, weight, bias, fp8)


def _conv_stem(w, chunks: torch.Tensor, emulate_bf16: bool) -> torch.Tensor:
    """[n, 1, 128, W] -> [n, 480, 16, W''] (:729-734), exact-erf GELU after each conv."""
    x = chunks
    for i in (1, 2, 3):
        x = F.conv2d(x, w[f"conv2d{i}.weight"], w[f"conv2d{i}.bias"], stride=2, padding=1)
        x = _rnd(F.gelu(_rnd(x, emulate_bf16)), emulate_bf16)
    return x


def _attention(w, prefix: str, h: torch.Tensor, wins, cfg: EncoderConfig, emulate_bf16: bool, fp8=None) -> torch.Tensor:
    hd, nh = cfg.head_dim, cfg.heads
    q = _rnd(_linear(h, w[prefix + "q_proj.weight"], w[prefix + "q_proj.bias"], fp8), emulate_bf16)
    k = _rnd(_linear(h, w[prefix + "k_proj.weight"], w[prefix + "k_proj.bias"], fp8), emulate_bf16)
    v = _rnd(_linear(h, w[prefix + "v_proj.weight"], w[prefix + "v_proj.bias"], fp8), emulate_bf16)
    out = torch.empty_like(q)
    s = 0
    scale = hd ** -0.5
    for wl in wins:
        qs = q[s : s + wl].view(wl, nh, hd).transpose(0, 1)
        ks = k[s : s + wl].view(wl, nh, hd).transpose(0, 1)
        vs = v[s : s + wl].view(wl, nh, hd).transpose(0, 1)
        att = torch.softmax(qs @ ks.transpose(1, 2) * scale, dim=-1, dtype=torch.float32)
        att = _rnd(att, emulate_bf16)
        out[s : s + wl] = (att @ vs).transpose(0, 1).reshape(wl, nh * hd)
        s += wl
    out = _rnd(out, emulate_bf16)
    return _rnd(_linear(out, w[prefix + "out_proj.weight"], w[prefix + "out_proj.bias"], fp8), emulate_bf16)


class oracle_device_fp32:
    """Context manager: true float32 matmuls / convolutions on a CUDA device (TF32 off) while the oracle runs there."""

    def __enter__(self):
        self._m, self._c = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        self._p = torch.get_float32_matmul_precision()
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self._m, self._c
        torch.set_float32_matmul_precision(self._p)
        return False


@torch.no_grad()
def encoder_forward(w, cfg: EncoderConfig, mels, emulate_bf16: bool = False, return_intermediate: bool = False, fp8=None, device=None):
    """mels: list of float32 [128, T_i] (already at the precision the tower sees,
    e.g. rounded through bf16).  Returns (hidden [sum tokens, output_dim] float32,
    token_lens list).  ``emulate_bf16`` rounds every module output through bf16,
    the reference deployment's rounding points.
    ``fp8``: None, "per_tensor" or "per_row" -- every nn.Linear (conv_out, q/k/v/out_proj, fc1, fc2, proj1, proj2)
    through ``linear_fp8``; the convolutions stay as they are (torchao's default filter).
    Per-tensor activation scales are taken over the whole call, as torchao takes them over the tensor it is handed.
    ``device``: where the same float32 torch ops run (default: where the weights are, normally the CPU).  On a CUDA device the
    caller must have TF32 switched off (``oracle_device_fp32``) -- it is the same restatement, only evaluated faster, so that
    whole batches at the real model sizes can be checked."""
    if fp8 is not None:
        emulate_bf16 = True
    if device is None:
        device = next(iter(w.values())).device
    device = torch.device(device)
    mels = [torch.as_tensor(np.asarray(m.cpu() if isinstance(m, torch.Tensor) else m), dtype=torch.float32).to(device) for m in mels]
    cf = cfg.chunk_frames
    d = cfg.d_model
    pe = sinusoid_table(cfg.max_source_positions, d).to(device)
    pe = _rnd(pe, emulate_bf16)

    # --- conv stem per chunk (all 100-frame-padded chunks batched; short lone clips alone)
    per_clip_tokens = []
    full_chunks, full_owner = [], []
    for ci, m in enumerate(mels):
        t = m.shape[1]
        toks = []
        for (s0, valid, padded) in chunk_plan(t, cf):
            ch = torch.zeros(cfg.num_mel_bins, padded, device=device)
            ch[:, :valid] = m[:, s0 : s0 + valid]
            if padded == cf:
                full_owner.append((ci, len(toks), valid))
                full_chunks.append(ch)
                toks.append(None)
            else:
                y = _conv_stem(w, ch[None, None], emulate_bf16)[0]  # [480, 16, w3]
                toks.append((y, valid))
        per_clip_tokens.append(toks)
    if full_chunks:
        ys = []
        batch = torch.stack(full_chunks)[:, None]
        for sl in batch.split(64, dim=0):
            ys.append(_conv_stem(w, sl, emulate_bf16))
        ys = torch.cat(ys, dim=0)
        for (ci, slot, valid), y in zip(full_owner, ys):
            per_clip_tokens[ci][slot] = (y, valid)

    # conv_out is ONE nn.Linear call over every chunk's time steps (:735-737) -- for the fp8 per-tensor variant its
    # dynamic activation scale therefore spans all of them, padded tail positions included
    emb_in, emb_meta = [], []
    for toks in per_clip_tokens:
        for (y, valid) in toks:
            c, f, tw = y.shape
            emb_in.append(y.permute(2, 0, 1).reshape(tw, c * f))  # index c*16+f (:735-736)
            emb_meta.append((tw, valid))
    emb_all = _rnd(_linear(torch.cat(emb_in, dim=0), w["conv_out.weight"], None, fp8), emulate_bf16) if emb_in else torch.zeros(0, d, device=device)
    rows, token_lens = [], []
    pos, mi = 0, 0
    for toks in per_clip_tokens:
        n_tok = 0
        for _ in toks:
            tw, valid = emb_meta[mi]
            mi += 1
            emb = _rnd(emb_all[pos : pos + tw] + pe[:tw], emulate_bf16)  # positions restart per chunk (:738-743)
            pos += tw
            nv = token_len(valid)
            rows.append(emb[:nv])
            n_tok += nv
        token_lens.append(n_tok)
    x = torch.cat(rows, dim=0) if rows else torch.zeros(0, d, device=device)
    inter = {"embed": x.clone()} if return_intermediate else None

    wins = []
    for m, n_tok in zip(mels, token_lens):
        wins += window_lens(n_tok, cfg, m.shape[1])

    for li in range(cfg.layers):
        p = f"layers.{li}."
        h = _rnd(F.layer_norm(x, (d,), w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"], 1e-5), emulate_bf16)
        x = _rnd(x + _attention(w, p + "self_attn.", h, wins, cfg, emulate_bf16, fp8), emulate_bf16)
        h = _rnd(F.layer_norm(x, (d,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], 1e-5), emulate_bf16)
        h = _rnd(_linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"], fp8), emulate_bf16)
        h = _rnd(F.gelu(h), emulate_bf16)
        h = _rnd(_linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"], fp8), emulate_bf16)
        x = _rnd(x + h, emulate_bf16)
        if return_intermediate and li == 0:
            inter["layer0"] = x.clone()

    x = _rnd(F.layer_norm(x, (d,), w["ln_post.weight"], w["ln_post.bias"], 1e-5), emulate_bf16)
    x = _rnd(_linear(x, w["proj1.weight"], w["proj1.bias"], fp8), emulate_bf16)
    x = _rnd(F.gelu(x), emulate_bf16)
    x = _rnd(_linear(x, w["proj2.weight"], w["proj2.bias"], fp8), emulate_bf16)
    if return_intermediate:
        return x, token_lens, inter
    return x, token_lens
