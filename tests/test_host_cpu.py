"""CPU-only checks of the product's host side: the C-ABI library loads without a GPU and exports every
symbol include/qasr_b200.h declares; host-only entry points agree with the golden vectors; the log-mel
kernel's __host__ __device__ math (FFT butterflies, real-input split, filter tables), compiled for the host,
agrees with the oracle; clip sharding is a valid deterministic partition."""

import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from qwen3_asr_b200.build import build

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    build()
    from qwen3_asr_b200 import load_library

    return load_library()


def test_library_exports_every_declared_symbol(lib):
    from qwen3_asr_b200._lib import SIGNATURES

    hdr = open(os.path.join(ROOT, "include", "qasr_b200.h")).read()
    declared = set(re.findall(r"QASR_API\s+[\w\s\*]+?\b(qasr_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(SIGNATURES), declared ^ set(SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.qasr_abi_version() == 1


def test_token_len_matches_transformers_golden(lib, golden):
    for t, n in zip(golden["toklen_T"], golden["toklen"]):
        assert lib.qasr_token_len(int(t)) == int(n), int(t)


def test_no_cpu_fallback_without_gpu():
    import torch

    from qwen3_asr_b200 import B200AudioEncoder, QasrError

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(QasrError):
        B200AudioEncoder({}, {})


def test_missing_library_fails_loudly(tmp_path):
    from qwen3_asr_b200 import QasrError, load_library

    with pytest.raises(QasrError):
        load_library(str(tmp_path / "nope.so"))


def test_create_rejects_bad_config_before_touching_cuda(lib):
    import ctypes as C

    from qwen3_asr_b200._lib import QasrConfig

    cfg = QasrConfig(d_model=1000, encoder_layers=2, encoder_attention_heads=16, encoder_ffn_dim=4096, output_dim=2048, n_window=50,
                     n_window_infer=800, downsample_hidden_size=480, num_mel_bins=128, max_source_positions=1500)
    h = C.c_void_p()
    assert lib.qasr_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"head_dim" in lib.qasr_last_error()


@pytest.fixture(scope="module")
def mel_host_bin(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("melhost") / "mel_host_test")
    src = os.path.join(ROOT, "tests", "host", "mel_host_test.cu")
    mel = os.path.join(ROOT, "qwen3_asr_b200", "csrc", "mel.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, src, mel],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("name", ["speech2s", "odd", "tone", "english_01"])
def test_mel_kernel_math_on_host(mel_host_bin, golden, tmp_path, name):
    """The kernel's own butterfly / split / filter code, executed on the CPU, against the f64 oracle."""
    from oracle.logmel import log10_mel_unclamped
    from test_oracle_golden import _clip

    x = _clip(golden, name)
    pcm, out = str(tmp_path / "pcm.bin"), str(tmp_path / "mel.bin")
    x.astype(np.float32).tofile(pcm)
    r = subprocess.run([mel_host_bin, pcm, out], capture_output=True, text=True)
    assert r.returncode == 0
    t = x.shape[0] // 160
    got = np.fromfile(out, dtype=np.float32).reshape(128, t)
    ref = log10_mel_unclamped(x)
    got = (np.maximum(got, got.max() - 8.0) + 4.0) / 4.0
    ref = (np.maximum(ref, ref.max() - 8.0) + 4.0) / 4.0
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-4
    assert np.abs(got - golden[f"mel_{name}"]).max() / np.abs(ref).max() <= 1e-4


def test_lpt_assign_is_a_balanced_partition():
    from qwen3_asr_b200.synth import lpt_assign

    rng = np.random.default_rng(0)
    frames = rng.integers(100, 3000, size=230).tolist()
    for n in (1, 2, 4, 8):
        bins = lpt_assign(frames, n)
        assert sorted(i for b in bins for i in b) == list(range(len(frames)))
        loads = [sum(frames[i] for i in b) for b in bins]
        assert max(loads) - min(loads) <= max(frames)
        assert bins == lpt_assign(frames, n)


def test_fast_gelu_formula_over_all_bf16_inputs():
    """The branch-free erf GELU of csrc/common.cuh (Abramowitz-Stegun 7.1.26), restated in float32 numpy,
    over every finite bf16 input: error far below a bf16 ulp, as good as torch's own float32 erf path."""
    import torch

    bits = np.arange(65536, dtype=np.uint32) << 16
    x = bits.view(np.float32)
    x = x[np.isfinite(x) & (np.abs(x) < 1e4)]
    f = np.float32
    a = np.abs(x)
    t = (f(1) / (f(0.231641888) * a + f(1))).astype(f)
    p = f(0.5307027145) * t + f(-0.7265760135)
    p = p * t + f(0.7107068705)
    p = p * t + f(-0.142248368)
    p = p * t + f(0.127414796)
    e = np.exp2((a * a * f(-0.72134752044)).astype(f)).astype(f)
    q = (p * t * e).astype(f)
    got = np.where(x >= 0, x - x * q, x * q).astype(f)
    xt = torch.from_numpy(x)
    ref64 = torch.nn.functional.gelu(xt.double())
    assert (torch.from_numpy(got).double() - ref64).abs().max().item() < 1e-6
    ref_bf = ref64.float().to(torch.bfloat16)
    mism = (torch.from_numpy(got).to(torch.bfloat16) != ref_bf).sum().item()
    torch_mism = (torch.nn.functional.gelu(xt).to(torch.bfloat16) != ref_bf).sum().item()
    assert mism <= 2 * max(torch_mism, 100), (mism, torch_mism)


def test_gelu_table_is_the_correctly_rounded_erf_gelu(lib):
    """qasr_gelu_table (host-only): bf16(x * ratio) over EVERY finite bf16 input equals the float64 erf GELU rounded to bf16 -- what
    the device evaluates in conv1 (one shared-memory load and one multiply per value).  torch's own float32 path misses that on
    ~129 inputs by one ulp, the closed-form approximation kept for the GEMM epilogues on 169."""
    import ctypes as C

    import torch

    buf = np.zeros(4096, dtype=np.float32)
    n = lib.qasr_gelu_table(buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size)
    assert n == 2 * 1666
    tab = buf[:n].reshape(2, 1666)
    bits = np.arange(65536, dtype=np.uint32)
    x = (bits << 16).view(np.float32)
    fin = np.isfinite(x)
    a = bits & 0x7FFF
    idx = np.clip(a, 0x3B00 - 1, 0x417F + 1) - (0x3B00 - 1)
    with np.errstate(invalid="ignore"):
        y = (x * tab[bits >> 15, idx]).astype(np.float32)
    got = torch.from_numpy(y[fin]).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(torch.from_numpy(x[fin]).double()).to(torch.bfloat16)   # float64 -> bf16, one rounding
    assert torch.equal(got.float(), ref.float())
    # sentinels: tiny inputs halve exactly, x >= 16 passes through, x <= -16 gives -0
    assert tab[0, 0] == 0.5 and tab[1, 0] == 0.5 and tab[0, -1] == 1.0 and tab[1, -1] == 0.0


def test_logmel_work_list_invariants(lib):
    """qasr_mel_plan (host-only): the list logmel_kernel_v3 walks.  Every 32-frame tile is transformed exactly once, in clip order, and
    clamped exactly once; a clamp rides at least `lag` = 8 x SMs items behind the LAST frame item of its clip (everything it waits for
    has a smaller ticket: no deadlock, and the clip maximum is final); `need` counts the clip's frame items; empty clips vanish."""
    import ctypes as C

    rng = np.random.default_rng(7)
    for n_sms, lens in ((148, [480000] * 32), (148, [0, 201, 80000, 0, 16000 * 30 + 77, 320, 7200]), (4, list(rng.integers(161, 200000, 40))),
                        (148, [16000 * 3600]), (2, [])):
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        n = lib.qasr_mel_plan(offs.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), n_sms, None, 0)
        buf = np.zeros((max(n, 1), 8), dtype=np.int64)
        assert lib.qasr_mel_plan(offs.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), n_sms, buf.ctypes.data_as(C.POINTER(C.c_int64)), n) == n
        items = buf[:n]
        frames = [int(x) // 160 for x in lens]
        tiles = [(c, f0, min(32, t - f0)) for c, t in enumerate(frames) for f0 in range(0, t, 32)]
        fr = [(int(i[0]), int(i[1]), int(i[2])) for i in items if i[2] > 0]
        assert fr == tiles                                                     # transformed once, in order
        cl = sorted((int(i[4]), int(i[5]), int(i[6])) for i in items if i[6] > 0)
        assert cl == sorted(tiles)                                             # clamped once
        col0 = np.concatenate([[0], np.cumsum(frames)])
        last = {}
        for k, i in enumerate(items):
            if i[2] > 0:
                last[int(i[0])] = k
                assert i[3] == col0[int(i[0])]
        lag = 8 * n_sms
        n_frame_items = len(tiles)
        for k, i in enumerate(items):
            if i[6] > 0:
                c = int(i[4])
                assert i[7] == (frames[c] + 31) // 32                          # need = frame items of the clip
                assert k > last[c]                                             # its clip's last frame item has a smaller ticket ...
                assert k >= min(last[c] + lag, n_frame_items)                  # ... by at least the lag, or it rides at the tail
        assert n <= 2 * len(tiles)


def test_pool_sharding_rule_on_the_host(lib):
    """qasr_pool_plan = the rule qasr_pool_submit shards by: contiguous clip ranges, near-equal mel-frame counts, every clip placed,
    empty ranges only when there are fewer clips than devices.  Pure host code: runs without a GPU."""
    import ctypes as C

    rng = np.random.default_rng(1234)

    def plan(lens, g):
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = np.full(len(lens), -1, dtype=np.int32)
        rc = lib.qasr_pool_plan(offs.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), g, out.ctypes.data_as(C.POINTER(C.c_int32)))
        assert rc == 0
        return out

    # BASELINE config 4: one hour of 1-30 s segments over 2 / 4 / 8 GPUs
    lens, total = [], 0
    while total < 3600 * 16000:
        ln = min(int(round(rng.uniform(1.0, 30.0) / 0.01)) * 160, 3600 * 16000 - total)
        lens.append(ln)
        total += ln
    frames = np.array(lens) // 160
    for g in (1, 2, 4, 8):
        s = plan(lens, g)
        assert s.min() == 0 and s.max() == g - 1 and np.all(np.diff(s) >= 0)            # contiguous, every device used
        load = np.array([frames[s == k].sum() for k in range(g)])
        assert load.max() <= frames.sum() / g + frames.max()                            # within one clip of the ideal share
        assert load.max() / (frames.sum() / g) - 1.0 <= 0.05                            # C4: < 5 % imbalance even at 8 GPUs
    # fewer clips than devices, equal clips, a zero-length clip, nothing at all
    assert plan([16000, 16000], 4).tolist() == [0, 2]                                  # one clip each on two of the four devices
    s = plan([16000] * 8, 4)
    assert s.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    s = plan([16000, 0, 16000, 16000], 2)
    assert np.all(np.diff(s) >= 0) and set(s.tolist()) == {0, 1}
    assert plan([], 3).tolist() == []


def test_pool_lpt_rule_matches_the_survey_definition(lib):
    """qasr_pool_plan_mode(LPT) = SURVEY 8(e): sort by mel frames descending, each clip to the least-loaded GPU -- the same rule as
    synth.lpt_assign (which bench.py --config c4 shards by); AUTO switches to it below four clips per GPU."""
    import ctypes as C

    from qwen3_asr_b200.synth import lpt_assign, workload_c4_lengths

    def plan(lens, g, mode):
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = np.full(len(lens), -1, dtype=np.int32)
        assert lib.qasr_pool_plan_mode(offs.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), g, mode, out.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        return out

    AUTO, CONTIG, LPT = 0, 1, 2
    lens = workload_c4_lengths()
    frames = [n // 160 for n in lens]
    for g in (2, 4, 8):
        s = plan(lens, g, LPT)
        bins = lpt_assign(frames, g)
        assert [sorted(np.where(s == k)[0].tolist()) for k in range(g)] == bins
        load = np.array([sum(frames[i] for i in b) for b in bins])
        assert load.max() / (sum(frames) / g) - 1.0 <= 0.02          # <= 2 % imbalance at 8 GPUs with ~230 segments
        assert plan(lens, g, AUTO).tolist() == plan(lens, g, CONTIG).tolist()      # many clips per GPU: zero-copy contiguous ranges
    # few, unequal clips: contiguous ranges cannot balance, LPT can
    few = [30 * 16000, 29 * 16000, 2 * 16000, 2 * 16000, 1 * 16000]
    c, l = plan(few, 2, CONTIG), plan(few, 2, LPT)
    fr = np.array(few) // 160
    worst = lambda s: max(fr[s == k].sum() for k in range(2))
    assert worst(l) < worst(c) and plan(few, 2, AUTO).tolist() == l.tolist()
