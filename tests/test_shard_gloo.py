"""N > 1 host logic on the CPU: two gloo ranks shard a batch without overlap, and the reported rate is
all units / max-over-ranks time (bench.py's multi-GPU contract).  No GPU, no oracle."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))


def test_shard_range_partitions():
    from spb200.shard import shard_range
    for total in (0, 1, 7, 64, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(512, 3, 8) == (192, 256)      # BASELINE configs[3]: 512 images over 8 GPUs


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from spb200.shard import shard_batch, max_over_ranks, whole_job_rate
    images = torch.arange(10 * 4, dtype=torch.float32).reshape(10, 4)
    mine = shard_batch(images, rank, world)
    # every rank reports which rows it took; gathered on rank 0
    rows = [None] * world
    dist.all_gather_object(rows, mine[:, 0].tolist())
    t = 0.5 + 0.25 * rank                                # rank 1 is slower
    tmax = max_over_ranks(t)
    rate = whole_job_rate(len(mine), t)
    if rank == 0:
        q.put((rows, tmax, rate))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_shard_and_reduce():
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    rows, tmax, rate = q.get()
    assert rows[0] == [0.0, 4.0, 8.0, 12.0, 16.0] and rows[1] == [20.0, 24.0, 28.0, 32.0, 36.0]
    assert abs(tmax - 0.75) < 1e-12
    assert abs(rate - 10 / 0.75) < 1e-9
