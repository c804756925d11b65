"""Multi-GPU host logic on CPU: clip sharding across ranks with world_size-2 gloo.  The data path has no
collective (clips are independent); torch.distributed only carries the bench's barrier / max-over-ranks, and
here the gather used to check that the shards partition the work."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qwen3_asr_b200.synth import lpt_assign

    mine = lpt_assign(frames, world)[rank]
    load = torch.tensor([float(sum(frames[i] for i in mine))], dtype=torch.float64)
    loads = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(loads, load)
    counts = torch.zeros(len(frames), dtype=torch.int64)
    counts[mine] = 1
    dist.all_reduce(counts)
    # the bench's timing reduction: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put(([float(v) for v in loads], counts.tolist(), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_shards_partition_the_hour_across_two_ranks():
    rng = np.random.default_rng(1234)
    frames = [int(round(rng.uniform(1.0, 30.0) / 0.01)) for _ in range(230)]  # config-4 segment lengths, in frames
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    loads, counts, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counts == [1] * len(frames), "every segment on exactly one rank"
    assert abs(loads[0] - loads[1]) <= max(frames)
    assert abs(sum(loads) - sum(frames)) < 1e-6
    assert tmax == 11.0
