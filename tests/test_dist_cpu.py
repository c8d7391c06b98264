"""CPU, world_size 2 over gloo: the N>1 plumbing (batch sharding without a data-path collective,
max-over-ranks timing, summed units) that bench.py uses under torchrun."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from softpool_b200 import dist as spd
    r, lr, w = spd.init(backend="gloo")
    lo, hi = spd.shard_batch(65, r, w)
    spd.barrier()
    t = spd.max_over_ranks(1.0 + r)            # slowest rank defines the step time
    units = spd.sum_over_ranks(hi - lo)        # whole-job units = sum of the shards
    assert spd.gather_over_ranks(10.0 + r) == [10.0, 11.0]      # per-rank step times, reported as min / median / max
    q.put((r, lo, hi, t, units))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 33), (33, 65)]      # contiguous, covers the batch once
    assert all(r[3] == 2.0 for r in res)
    assert all(r[4] == 65.0 for r in res)


def test_shard_batch_covers_everything():
    from softpool_b200.dist import shard_batch
    for gb in (0, 1, 7, 32, 256):
        for w in (1, 2, 3, 4, 8):
            cuts = [shard_batch(gb, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == gb
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
