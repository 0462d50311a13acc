"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous track shards cover the batch exactly once,
the host-side gather restores batch order, and the bench's timing reduction is a max over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import stratum_dsp_b200 as S


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1024, 8192, 8191):
        for nd in (1, 2, 3, 4, 8):
            b = [S.shard_bounds(n, d, nd) for d in range(nd)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(nd - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_tracks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = S.shard_bounds(n_tracks, rank, world)
    local = [(i, 70.0 + (i % 111)) for i in range(a, b)]  # (track id, "bpm") stand-ins for result structs
    gathered = [None] * world
    dist.all_gather_object(gathered, local)  # host-side gather of fixed-size results; no data-path collective
    t = torch.tensor([10.0 + rank * 5.0], dtype=torch.float64)  # per-rank step time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        flat = [x for part in gathered for x in part]
        q.put((flat, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, tmax = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [i for i, _ in flat] == list(range(n))  # batch order restored, every track exactly once
    assert tmax == 15.0  # max over ranks
