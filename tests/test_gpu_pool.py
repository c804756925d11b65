"""qasr_pool_*: one process, several GPUs (or several handles on one GPU) -- sharding must not change a single bit."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clips(n, seed0=50):
    from oracle.signals import speech_like

    rng = np.random.default_rng(7)
    return [speech_like(int(rng.integers(30, 1200)) * 160 + int(rng.integers(0, 160)), seed0 + i) for i in range(n)]


@pytest.fixture(scope="module")
def tiny_cfg():
    from oracle import CONFIGS, make_weights

    cfg = CONFIGS["tiny"]
    return cfg, make_weights(cfg, seed=1)


def _pack(clips):
    offs = np.zeros(len(clips) + 1, dtype=np.int64)
    for i, c in enumerate(clips):
        offs[i + 1] = offs[i] + c.shape[0]
    pcm = torch.empty(int(offs[-1]), dtype=torch.float32, pin_memory=True)
    for i, c in enumerate(clips):
        pcm[int(offs[i]):int(offs[i + 1])] = torch.from_numpy(c)
    return pcm, offs


@pytest.mark.parametrize("n_handles", [1, 2, 3])
def test_pool_matches_single_handle_bitwise(tiny_cfg, n_handles):
    from qwen3_asr_b200 import B200AudioEncoder, B200EncoderPool

    cfg, w = tiny_cfg
    n_dev = torch.cuda.device_count()
    devices = [i % n_dev for i in range(n_handles)]     # on a 1-GPU box: several handles on cuda:0 (same code path, same threads)
    clips = _clips(11)
    enc = B200AudioEncoder(cfg, w, max_chunks=64)
    want, want_toks = enc.encode_pcm(clips)
    torch.cuda.synchronize()
    want = want.cpu()
    enc.close()

    pool = B200EncoderPool(cfg, w, devices=devices, max_chunks=64, sharding="contiguous")
    assert len(pool) == n_handles
    pcm, offs = _pack(clips)
    out = torch.zeros((int(want_toks.sum()), pool.output_dim), dtype=torch.bfloat16, pin_memory=True)
    ticket, toks, devs = pool.submit_pcm_host(pcm, offs, out)
    pool.collect(ticket)
    assert toks.tolist() == want_toks.tolist()
    assert torch.equal(out, want)                         # clip order, bit-exact: sharding is invisible
    # contiguous ranges, non-decreasing worker index, every worker used when there are enough clips
    assert set(devs.tolist()) <= set(devices)
    assert all(devices.index(a) <= devices.index(b) for a, b in zip(devs[:-1], devs[1:])) or n_dev < n_handles
    frames = np.array([c.shape[0] // 160 for c in clips])
    if n_dev >= n_handles:   # balance: no GPU carries more than the ideal share plus one (largest) clip
        load = [int(frames[devs == d].sum()) for d in devices]
        assert max(load) <= frames.sum() / n_handles + frames.max()
    pool.close()


def test_pool_balance_and_pipelining(tiny_cfg):
    from qwen3_asr_b200 import B200EncoderPool

    cfg, w = tiny_cfg
    n_dev = torch.cuda.device_count()
    devices = [i % n_dev for i in range(2)]
    pool = B200EncoderPool(cfg, w, devices=devices, max_chunks=64)
    batches = [_clips(8, seed0=100 + 10 * b) for b in range(4)]
    packed = [_pack(c) for c in batches]
    outs, tickets = [], []
    for (pcm, offs), clips in zip(packed, batches):
        total = int(sum(pool.lib.qasr_token_len(c.shape[0] // 160) for c in clips))
        out = torch.zeros((total, pool.output_dim), dtype=torch.bfloat16, pin_memory=True)
        t, toks, devs = pool.submit_pcm_host(pcm, offs, out)       # four batches in flight before the first collect
        outs.append(out)
        tickets.append(t)
    for t in reversed(tickets):                                      # collect out of order
        pool.collect(t)
    again, _ = pool.encode_pcm(batches[2])
    assert torch.equal(again, outs[2])                              # deterministic across calls and positions in the queue
    with pytest.raises(Exception):
        pool.collect(tickets[0])                                     # a ticket can be collected once
    pool.close()


def test_pool_error_reporting(tiny_cfg):
    from qwen3_asr_b200 import B200EncoderPool, QasrError

    cfg, w = tiny_cfg
    pool = B200EncoderPool(cfg, w, devices=[0], max_chunks=64)
    pcm, offs = _pack([np.zeros(100, np.float32), np.zeros(16000, np.float32)])   # first clip shorter than the reflect pad
    out = torch.zeros((64, pool.output_dim), dtype=torch.bfloat16, pin_memory=True)
    ticket, _, _ = pool.submit_pcm_host(pcm, offs, out)
    with pytest.raises(QasrError, match="200 samples"):
        pool.collect(ticket)
    small = torch.zeros((1, pool.output_dim), dtype=torch.bfloat16, pin_memory=True)
    pcm2, offs2 = _pack(_clips(2))
    with pytest.raises(QasrError, match="too small"):
        pool.submit_pcm_host(pcm2, offs2, small)
    pool.close()
