"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: stats-block reduction and the
read-range / contig sharding helpers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pbsim_b200 import stats_reduce as SR


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cells = 16 + 100001 + 202
    t = torch.zeros(cells, dtype=torch.int64)
    t[0] = 10 + rank           # reads
    t[2] = 1000 * (rank + 1)   # bases
    t[SR.CELL_LEN_MIN] = 100 + 5 * rank
    t[SR.CELL_LEN_MAX] = 900 - 7 * rank
    t[16 + 85000 + rank] = 3   # freq_accuracy
    t[16 + 100001 + 150] = 2   # freq_len
    SR.reduce_stats_tensor(t, dist)
    start, total = SR.emitted_prefix(1000 * (rank + 1), dist)
    q.put((rank, t[0].item(), t[2].item(), t[SR.CELL_LEN_MIN].item(), t[SR.CELL_LEN_MAX].item(),
           t[16 + 85000].item(), t[16 + 85001].item(), t[16 + 100001 + 150].item(), start, total))
    dist.destroy_process_group()


def test_stats_reduce_and_prefix_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, reads, bases, mn, mx, fa0, fa1, fl, start, total in res:
        assert (reads, bases, mn, mx, fa0, fa1, fl) == (21, 3000, 100, 900, 3, 3, 4)
        assert total == 3000 and start == (0 if rank == 0 else 1000)


def test_sharding_helpers_cover_everything_once():
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            seen += SR.contigs_for_rank(24, r, world)
        assert sorted(seen) == list(range(24))
        for n in (0, 1, 7, 1000, 1001):
            ranges = [SR.read_range_for_rank(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
