"""The BASELINE.json configs that round 1 never ran on hardware, as GPU tests through the C ABI:
C5(ii) dual-model residency with the C3 routing, a 0.6B 32 x 30 s batch, and C4 through the one-process pool."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HID_TOL = 2e-2


def _range_rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max())


def test_dual_model_residency_with_c3_routing():
    """BASELINE configs[4](ii) / src/server.py:411-425, 1345-1356: a 0.6B and a 1.7B handle resident on one GPU; WS partials
    (use_fast = not pad_silence) go to the 0.6B backend, flushes (+600 ms silence) to the 1.7B one.  Each model's batch is checked
    against the oracle of ITS weights, and neither handle disturbs the other (interleaved calls, bit-identical repeats)."""
    from oracle import CONFIGS, encoder_forward, make_weights, oracle_device_fp32
    from qwen3_asr_b200 import B200AudioEncoder, B200PreFrontend
    from qwen3_asr_b200.synth import workload_c3

    free0, _ = torch.cuda.mem_get_info()
    cfgs = {n: CONFIGS[n] for n in ("0.6B", "1.7B")}
    ws = {n: make_weights(cfgs[n], seed=21 + i) for i, n in enumerate(cfgs)}
    encs = {n: B200AudioEncoder(cfgs[n], ws[n], max_chunks=256) for n in cfgs}
    try:
        free1, _ = torch.cuda.mem_get_info()
        held = sum(e.device_bytes for e in encs.values())
        print(f"\nboth models resident: {held / 2**30:.2f} GiB held by the two handles ({(free0 - free1) / 2**30:.2f} GiB less free memory); "
              f"weights {encs['0.6B'].weight_bytes / 2**20:.0f} + {encs['1.7B'].weight_bytes / 2**20:.0f} MiB")
        wins, flush = workload_c3(n_streams=32)
        route = {"0.6B": [i for i, f in enumerate(flush) if not f], "1.7B": [i for i, f in enumerate(flush) if f]}
        pres = {n: B200PreFrontend(encs[n]) for n in encs}
        outs = {}
        for rep in range(2):                      # interleave the two handles, twice
            for n in ("0.6B", "1.7B"):
                hid, toks = pres[n].encode_windows([wins[i] for i in route[n]], 16000, pad_silence=[flush[i] for i in route[n]])
                torch.cuda.synchronize()
                if rep == 0:
                    outs[n] = (hid.clone(), toks)
                else:
                    assert torch.equal(hid, outs[n][0]), f"{n}: a repeat differs after the other model ran"
        for n in encs:
            hid, toks = outs[n]
            assert hid.shape[1] == cfgs[n].output_dim
            pcm, offs = pres[n].prepare([wins[i] for i in route[n]], 16000, pad_silence=[flush[i] for i in route[n]])
            mel, flens = encs[n].logmel_packed(pcm, offs)
            cols = np.concatenate([[0], np.cumsum(flens)])
            mels = [mel[:, cols[i]:cols[i + 1]].to(torch.bfloat16).float() for i in range(len(flens))]
            w_dev = {k: v.cuda() for k, v in ws[n].items()}
            with oracle_device_fp32():
                ref, rt = encoder_forward(w_dev, cfgs[n], mels, device="cuda")
            assert list(rt) == toks.tolist()
            err = _range_rel(hid.float(), ref)
            print(f"  {n}: {len(route[n])} windows, {int(toks.sum())} tokens, range-rel error vs the fp32 oracle {err:.2e}")
            assert err <= HID_TOL
    finally:
        for e in encs.values():
            e.close()


def test_06b_full_batch_is_batch_invariant_and_matches_the_oracle():
    """A 0.6B 32 x 30 s batch (d = 896 = 3.5 x 256: the N tile falls back to 128): three clips bit-identical to lone runs, one
    clip against the fp32 oracle."""
    from oracle import CONFIGS, encoder_forward, make_weights, oracle_device_fp32
    from oracle.signals import speech_like
    from qwen3_asr_b200 import B200AudioEncoder

    cfg = CONFIGS["0.6B"]
    w = make_weights(cfg, seed=3)
    enc = B200AudioEncoder(cfg, w)
    try:
        clips = [speech_like(30 * 16000, 300 + i) for i in range(32)]
        out, toks = enc.encode_pcm(clips)
        torch.cuda.synchronize()
        assert toks.tolist() == [390] * 32
        for i in (0, 13, 31):
            alone, _ = enc.encode_pcm([clips[i]])
            torch.cuda.synchronize()
            assert torch.equal(alone, out[390 * i:390 * (i + 1)])
        mel, flens = enc.logmel([clips[13]])
        w_dev = {k: v.cuda() for k, v in w.items()}
        with oracle_device_fp32():
            ref, _ = encoder_forward(w_dev, cfg, [mel.to(torch.bfloat16).float()], device="cuda")
        assert _range_rel(out[390 * 13:390 * 14].float(), ref) <= HID_TOL
    finally:
        enc.close()


def test_c4_hour_through_the_pool_matches_one_handle():
    """BASELINE configs[3] through B200EncoderPool (one process, every visible GPU -- with one GPU, several handles on it): the pool's
    output is bit-identical to a single handle's, in clip order, in both sharding modes (contiguous ranges and LPT)."""
    from qwen3_asr_b200 import B200AudioEncoder, B200EncoderPool
    from qwen3_asr_b200.synth import model_config, random_weights, workload_c4_clips, workload_c4_lengths

    cfg = model_config("0.6B")       # the sharding logic is model-independent: the smaller model keeps the test short
    w = random_weights(cfg, seed=0)
    lens = workload_c4_lengths(total_seconds=600)
    clips = workload_c4_clips(lens, range(len(lens)))
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0, 0]
    enc = B200AudioEncoder(cfg, w, max_chunks=512)
    try:
        ref, toks = enc.encode_pcm(clips)
        torch.cuda.synchronize()
        ref = ref.cpu()
    finally:
        enc.close()
    for mode in ("contiguous", "lpt"):
        pool = B200EncoderPool(cfg, w, devices=devices, max_chunks=512, sharding=mode)
        try:
            out, ptoks = pool.encode_pcm(clips)
            assert ptoks.tolist() == toks.tolist()
            assert torch.equal(out, ref), f"pool ({mode}) differs from a single handle"
            # few clips per device: LPT balances where contiguous ranges cannot
            few = [clips[i] for i in np.argsort(lens)[-len(devices) - 1:]]
            o2, t2 = pool.encode_pcm(few)
            offs = np.concatenate([[0], np.cumsum(t2)])
            for j, c in enumerate(few):
                k = int(np.where([len(c) == len(x) and np.array_equal(c, x) for x in clips])[0][0])
                r0 = int(np.sum(toks[:k]))
                assert torch.equal(o2[offs[j]:offs[j + 1]], ref[r0:r0 + int(toks[k])])
        finally:
            pool.close()
