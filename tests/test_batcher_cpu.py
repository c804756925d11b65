"""WindowBatcher host logic with a stub encoder -- CPU only."""

import threading
import time

import numpy as np
import pytest

from qwen3_asr_b200.batcher import WindowBatcher


def _stub(calls):
    def encode(windows, flush):
        calls.append((len(windows), list(flush)))
        lens = [max(1, len(w) // 1600) for w in windows]
        rows = [np.full((n, 4), float(w[0]), dtype=np.float32) for w, n in zip(windows, lens)]
        return np.concatenate(rows, 0), np.array(lens)
    return encode


def test_concurrent_windows_become_one_batch_and_results_route_back():
    calls = []
    b = WindowBatcher(_stub(calls), max_wait_ms=150.0)
    wins = [np.full(1600 * (i + 1), i + 1, dtype=np.float32) for i in range(8)]
    futs = [None] * 8

    def client(i):
        futs[i] = b.submit(wins[i], flush=(i % 4 == 3))
    ts = [threading.Thread(target=client, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    outs = [f.result(timeout=5) for f in futs]
    b.close()
    assert sum(c[0] for c in calls) == 8 and len(calls) <= 2          # gathered, not one call per window
    for i, o in enumerate(outs):
        assert o.shape == (i + 1, 4) and (o == i + 1).all()            # every client got its own clip's rows


def test_flushes_on_audio_budget_without_waiting():
    calls = []
    b = WindowBatcher(_stub(calls), max_wait_ms=10_000.0, max_audio_s=1.0)
    t0 = time.monotonic()
    futs = [b.submit(np.ones(8000, np.float32)) for _ in range(4)]     # 4 x 0.5 s: the budget splits them into pairs
    [f.result(timeout=5) for f in futs]
    assert time.monotonic() - t0 < 5.0
    b.close()
    assert [c[0] for c in calls] == [2, 2]


def test_failure_reaches_every_future_of_the_batch_and_close_drains():
    def boom(windows, flush):
        raise ValueError("encoder failed")
    b = WindowBatcher(boom, max_wait_ms=20.0)
    futs = [b.submit(np.ones(1600, np.float32)) for _ in range(3)]
    for f in futs:
        with pytest.raises(ValueError):
            f.result(timeout=5)
    b.close()
    with pytest.raises(RuntimeError):
        b.submit(np.ones(1600, np.float32))
    calls = []
    b2 = WindowBatcher(_stub(calls), max_wait_ms=10_000.0)
    f = b2.submit(np.ones(1600, np.float32))
    b2.close()                                                           # pending work is run, not dropped
    assert f.result(timeout=1).shape == (1, 4)
