"""CPU, world_size 2, gloo: the N>1 plumbing of the path (image shards + max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiposenet.pytorch_b200 import shard


def test_shard_range_partitions():
    for total in (0, 1, 7, 32, 33, 64):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the training step's single collective: flat gradient mean over ranks
    grads = [("a.weight", torch.full((2, 3), float(rank + 1))), ("b.bias", torch.arange(4.0) * (rank + 1))]
    flat, index = shard.flatten_grads(grads)
    shard.allreduce_mean_(flat)
    back = shard.unflatten_grads(flat, index)
    assert torch.allclose(back["a.weight"], torch.full((2, 3), 1.5)) and torch.allclose(back["b.bias"], torch.arange(4.0) * 1.5)
    b, e = shard.shard_range(33, rank, world)
    slow = shard.max_over_ranks(10.0 * (rank + 1))
    rate = shard.whole_job_rate(e - b, 10.0 * (rank + 1))
    out.put((rank, b, e, slow, rate))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_rate():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in ps]
    assert [r[1:3] for r in res] == [(0, 17), (17, 33)]
    assert all(abs(r[3] - 20.0) < 1e-9 for r in res)
    assert all(abs(r[4] - 33 / 0.020) < 1e-6 for r in res)  # all images over the slowest rank's time
