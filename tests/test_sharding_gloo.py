"""N>1 host logic on CPU: world_size-2 `gloo` run of the stream sharding + max-over-ranks timing
reduce that bench.py uses (no data-path collective exists on this path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glava_b200.shard import shard_streams, max_over_ranks, gather_objects
from glava_b200.synth import synth_batch_int16


def test_partition_covers_every_stream_once():
    for total in (1, 2, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = shard_streams(total, world, r)
                seen += list(range(first, first + count))
            assert seen == list(range(total))
            counts = [shard_streams(total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_streams(total, world, rank)
    pcm = synth_batch_int16(first, count, 0, 64)
    digest = [(first + i, int(pcm[i].astype(np.int64).sum())) for i in range(count)]
    everyone = gather_objects(digest)
    slowest = max_over_ranks(10.0 + rank)
    dist.barrier()
    if rank == 0:
        q.put((sorted(sum(everyone, [])), slowest))
    dist.destroy_process_group()


def test_two_rank_shard_equals_single_process():
    total, world = 5, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    digest, slowest = q.get()
    ref = synth_batch_int16(0, total, 0, 64)
    assert digest == [(i, int(ref[i].astype(np.int64).sum())) for i in range(total)]
    assert slowest == 11.0                         # max over ranks, not rank 0's own time
