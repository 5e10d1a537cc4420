"""World-size-2 gloo test of the multi-GPU plumbing (no GPU needed): shards are disjoint and cover the job,
timings combine with MAX, counts with SUM — what bench.py does at N > 1."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from april_asr_b200 import dist_util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, l = dist_util.rank_info()
    ids = dist_util.shard_stream_ids(r, w, 4)
    dist.barrier()
    mx, sm = dist_util.combine(dist, "cpu", [0.1 * (r + 1), 5.0 - r], [len(ids), 100.0 * (r + 1)])
    out.put((r, ids, mx, sm))
    dist.destroy_process_group()


def test_sharding_and_reductions_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = res[0][1] + res[1][1]
    assert ids == list(range(8))                                   # disjoint, complete
    for _, _, mx, sm in res:
        assert abs(mx[0] - 0.2) < 1e-12 and abs(mx[1] - 5.0) < 1e-12   # max over ranks
        assert sm == [8.0, 300.0]                                  # sums


def test_single_rank_identity():
    assert dist_util.combine(None, "cpu", [1.5], [2.0]) == ([1.5], [2.0])
    assert dist_util.shard_stream_ids(3, 8, 64) == list(range(192, 256))
