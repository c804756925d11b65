"""Kernel-level parity tests on the B200: each CUDA kernel, called through the C ABI's bring-up
hooks, against a float32 PyTorch statement of the same op (floating-point kernels) on seeded inputs."""

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from qwen3_asr_b200 import load_library

    return load_library()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gemm(lib, a, b, bias, residual, act, impl):
    from qwen3_asr_b200._lib import check

    m, k = a.shape
    n = b.shape[0]
    d = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    check(lib, lib.qasr_debug_gemm(_ptr(a), _ptr(b), _ptr(bias), _ptr(residual), _ptr(d), m, n, k, act, impl, _stream()), "qasr_debug_gemm")
    torch.cuda.synchronize()
    return d


def _gemm_ref(a, b, bias, residual, act):
    y = a.float() @ b.float().t()
    if bias is not None:
        y = y + bias
    y = y.to(torch.bfloat16).float()
    if act == 1:
        y = torch.nn.functional.gelu(y).to(torch.bfloat16).float()
    if residual is not None:
        y = (y + residual.float())
    return y.to(torch.bfloat16).float()


GEMM_SHAPES = [
    # m, n, k, act, residual
    (128, 128, 64, 0, False),
    (128, 256, 128, 0, False),
    (300, 256, 1024, 1, False),
    (77, 192, 128, 0, True),
    (1000, 1024, 1024, 0, True),
    (513, 3072, 1024, 0, False),
    (390, 4096, 1024, 1, False),
    (390, 1024, 4096, 0, True),
    (260, 896, 896, 0, False),
    (260, 3584, 896, 1, False),
]


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_gemm(lib, shape, impl):
    m, n, k, act, use_res = shape
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(n, k, device="cuda", generator=g) / np.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda", generator=g) * 0.1
    res = torch.randn(m, n, device="cuda", generator=g).to(torch.bfloat16) if use_res else None
    got = _gemm(lib, a, b, bias, res, act, impl).float()
    ref = _gemm_ref(a, b, bias, res, act)
    assert torch.isfinite(got).all(), f"non-finite output: {(~torch.isfinite(got)).sum().item()} of {got.numel()}"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    # one bf16 ulp of the result scale: accumulation order differs from torch's, rounding can flip
    assert err <= 2 ** -7 * scale, f"max abs err {err} (scale {scale})"
    # the bulk must be exact or 1 ulp: catch a wrong-but-close kernel
    frac_off = ((got - ref).abs() > 2 ** -8 * ref.abs().clamp_min(1e-3)).float().mean().item()
    assert frac_off < 0.02, f"{frac_off:.4f} of elements off by more than 1 bf16 ulp"


def test_gemm_tc_matches_simt_bitwise_mostly(lib):
    """The two implementations share the epilogue; fp32 accumulation order differs only inside the K sum."""
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(640, 1024, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(1024, 1024, device="cuda", generator=g) / 32).to(torch.bfloat16)
    bias = torch.randn(1024, device="cuda", generator=g)
    y0 = _gemm(lib, a, b, bias, None, 0, 0).float()
    y1 = _gemm(lib, a, b, bias, None, 0, 1).float()
    assert (y0 != y1).float().mean().item() < 0.05
    assert (y0 - y1).abs().max().item() <= 2 ** -7 * y1.abs().max().item()


@pytest.mark.parametrize("rows,d", [(1, 128), (65, 896), (1000, 1024), (333, 2048)])
def test_layernorm(lib, rows, d):
    from qwen3_asr_b200._lib import check

    g = torch.Generator(device="cuda").manual_seed(rows + d)
    x = (torch.randn(rows, d, device="cuda", generator=g) * 3 + 0.5).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    out = torch.empty_like(x)
    check(lib, lib.qasr_debug_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), rows, d, _stream()), "layernorm")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (d,), gamma, beta, 1e-5)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item(), err


def _attention_ref(qkv, wins, d, heads):
    q, k, v = qkv.float().split(d, dim=1)
    out = torch.zeros_like(q)
    for s, wl in wins:
        qs = q[s:s + wl].view(wl, heads, 64).transpose(0, 1)
        ks = k[s:s + wl].view(wl, heads, 64).transpose(0, 1)
        vs = v[s:s + wl].view(wl, heads, 64).transpose(0, 1)
        att = torch.softmax(qs @ ks.transpose(1, 2) * 0.125, dim=-1)
        out[s:s + wl] = (att @ vs).transpose(0, 1).reshape(wl, d)
    return out


@pytest.mark.parametrize("impl", ["tcgen05", "mma_sync"])
@pytest.mark.parametrize("heads,lens", [(2, [6]), (2, [104, 33]), (16, [104, 104, 104, 78, 1, 13, 17]), (14, [104, 65, 104]), (2, [128, 127, 16, 15]),
                                        (16, [104] * 40 + [78] * 10)])
def test_window_attention(lib, heads, lens, impl):
    from qwen3_asr_b200._lib import check

    d = heads * 64
    n = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(n + heads)
    qkv = torch.randn(n, 3 * d, device="cuda", generator=g)
    qkv[:, :2 * d] *= 2.0  # peaky softmax
    qkv = qkv.to(torch.bfloat16)
    wins, s = [], 0
    for wl in lens:
        wins.append((s, wl))
        s += wl
    win_host = (C.c_int32 * (2 * len(wins)))(*[v for w in wins for v in w])
    out = torch.full((n, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    if impl == "tcgen05":
        check(lib, lib.qasr_debug_attention_tc(_ptr(qkv), _ptr(out), win_host, len(wins), n, d, heads, _stream()), "attention_tc")
    else:
        check(lib, lib.qasr_debug_attention(_ptr(qkv), _ptr(out), win_host, len(wins), d, heads, _stream()), "attention")
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, wins, d, heads)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    # P is rounded to bf16 before PV (as flash-attn does): ~2^-8 relative on a convex combination of |v| <~ 4
    assert err <= 0.03, err


# ---------------------------------------------------------------------------------------------- fp8 (QUANTIZE=fp8 variant)
def _quant_ref(x, per_row):
    """torchao's dynamic e4m3 recipe on the CPU: scale = max(amax, 1e-12) / 448, q = e4m3_rn_sat(x / scale)."""
    xf = x.float().cpu()
    amax = xf.abs().amax(dim=1, keepdim=True) if per_row else xf.abs().amax().reshape(1, 1).expand(xf.shape[0], 1)
    scale = amax.clamp(min=1e-12) / 448.0
    q = (xf / scale).clamp(-448, 448).to(torch.float8_e4m3fn)
    return q, scale.reshape(-1).contiguous()


@pytest.mark.parametrize("per_row", [0, 1])
@pytest.mark.parametrize("rows,k", [(5, 128), (300, 1024), (1000, 4096), (39, 7680)])
def test_quant_fp8_kernel_bit_exact(lib, rows, k, per_row):
    from qwen3_asr_b200._lib import check

    g = torch.Generator().manual_seed(rows * 7 + k + per_row)
    x = (torch.randn(rows, k, generator=g) * torch.logspace(-2, 1, rows)[:, None]).to(torch.bfloat16)
    x[0, :8] = 0
    if rows > 4:
        x[3] = 0  # an all-zero row: scale clamps to 1e-12 / 448, q = 0
    xd = x.cuda()
    q = torch.zeros((rows, k), dtype=torch.uint8, device="cuda")
    sc = torch.zeros(rows, dtype=torch.float32, device="cuda")
    check(lib, lib.qasr_debug_quant_fp8(_ptr(xd), rows, k, _ptr(q), _ptr(sc), per_row, _stream()), "qasr_debug_quant_fp8")
    q_ref, sc_ref = _quant_ref(x, bool(per_row))
    assert torch.equal(sc.cpu(), sc_ref)
    assert torch.equal(q.cpu(), q_ref.view(torch.uint8))


FP8_GEMM_SHAPES = [
    # m, n, k, act, residual
    (128, 128, 128, 0, False),
    (300, 256, 1024, 1, False),
    (77, 192, 256, 0, True),
    (1000, 1024, 1024, 0, True),
    (513, 3072, 1024, 0, False),
    (390, 1024, 4096, 0, True),
    (200, 896, 3584, 1, False),
    (260, 1024, 7680, 0, False),
]


@pytest.mark.parametrize("m,n,k,act,res", FP8_GEMM_SHAPES)
def test_gemm_fp8_tcgen05_vs_torch(lib, m, n, k, act, res):
    """e4m3 x e4m3 tcgen05 GEMM (kind::f8f6f4) with row / column scales against the dequantised fp32 product."""
    from qwen3_asr_b200._lib import check

    g = torch.Generator().manual_seed(m + n + k)
    a = (torch.randn(m, k, generator=g) * 3).clamp(-448, 448).to(torch.float8_e4m3fn)
    b = (torch.randn(n, k, generator=g) * 3).clamp(-448, 448).to(torch.float8_e4m3fn)
    rs = torch.rand(m, generator=g) * 0.02 + 0.001
    cs = torch.rand(n, generator=g) * 0.02 + 0.001
    bias = torch.randn(n, generator=g)
    resid = torch.randn(m, n, generator=g).to(torch.bfloat16) if res else None
    y = (a.float() @ b.float().t()) * rs[:, None] * cs[None, :] + bias
    y = y.to(torch.bfloat16).float()
    if act == 1:
        y = torch.nn.functional.gelu(y).to(torch.bfloat16).float()
    if res:
        y = (y + resid.float()).to(torch.bfloat16).float()
    d = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    rd = resid.cuda() if res else None
    ad, bd, rsd, csd, biasd = a.view(torch.uint8).cuda(), b.view(torch.uint8).cuda(), rs.cuda(), cs.cuda(), bias.cuda()  # keep alive
    check(lib, lib.qasr_debug_gemm_fp8(_ptr(ad), _ptr(bd), _ptr(rsd), _ptr(csd), _ptr(biasd), _ptr(rd), _ptr(d), m, n, k, act, _stream()),
          "qasr_debug_gemm_fp8")
    torch.cuda.synchronize()
    got = d.float().cpu()
    assert torch.isfinite(got).all()
    # products of e4m3 values are exact in fp32; only the accumulation order differs -> at most a bf16 ulp after rounding
    err = ((got - y).abs() / (y.abs() + 1e-2)).max().item()
    assert err <= 2 ** -7, err
