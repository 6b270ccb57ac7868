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


def test_lpt_split_covers_every_sequence_once_and_balances():
    sizes = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
    for world in (1, 2, 4, 8):
        parts = SR.lpt_assign(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert sum(sizes) / world / max(loads) > 0.95   # the 24 human-sized contigs balance to > 95 % on 8 ranks


def test_merged_summary_from_a_reduced_block():
    import numpy as np
    cells = np.zeros(16 + 100001 + 2002, dtype=np.int64)
    # two reads: lengths 100 and 300, accuracies 0.80 and 0.90
    cells[0] = 2; cells[1] = 2; cells[2] = 400; cells[3] = 100; cells[4] = 300
    cells[5], cells[6], cells[7] = 4, 8, 12
    cells[SR.CELL_ACC_FX] = int(round(0.8 * 2 ** 40)) + int(round(0.9 * 2 ** 40))
    cells[16 + 80000] = 1; cells[16 + 90000] = 1
    cells[16 + 100001 + 100] = 1; cells[16 + 100001 + 300] = 1
    m = SR.merged_summary(cells, 1000)
    assert m["res_num"] == 2 and m["res_len_mean"] == 200.0
    assert abs(m["res_accuracy_mean"] - 0.85) < 1e-12 and abs(m["res_accuracy_sd"] - 0.05) < 1e-9
    assert abs(m["res_len_sd"] - 100.0) < 1e-9


def _check_plan(reads_est, world, plan):
    # every sequence is covered exactly once by consecutive read ranges; the last part runs to the quota
    by_seq = {}
    for r, parts in enumerate(plan):
        for p in parts:
            by_seq.setdefault(p["seq"], []).append((p["first_read"], p["max_reads"], p["last"], r))
    assert sorted(by_seq) == list(range(len(reads_est)))
    for k, parts in by_seq.items():
        parts.sort()
        assert parts[0][0] == 0
        for a, b in zip(parts, parts[1:]):
            assert not a[2] and a[1] > 0 and a[0] + a[1] == b[0]      # feeders are bounded and contiguous
            assert a[3] <= b[3]                                       # line order follows rank order
        assert parts[-1][2] and parts[-1][1] == 0
        # a part that is not the last one ends safely in front of the estimated end of the sequence
        if len(parts) > 1:
            assert parts[-1][0] <= 0.961 * reads_est[k]
            assert parts[1][0] + 1 >= 0.0199 * reads_est[k]
            assert reads_est[k] >= 2000   # short sequences are never cut
    # ranks own contiguous pieces of the line
    flat = [(p["seq"], p["first_read"]) for parts in plan for p in parts]
    assert flat == sorted(flat)


def test_line_split_plans():
    sizes = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
    reads = [s * 1000.0 for s in sizes]
    for world in (1, 2, 3, 4, 8, 16, 40):
        for n in (1, 2, 5, 20, 24):
            plan = SR.plan_line_split(reads[:n], world)
            assert len(plan) == world
            _check_plan(reads[:n], world, plan)
            if world <= 8 and n >= 20:
                loads = [sum(p["est"] for p in parts) for parts in plan]
                assert min(loads) / max(loads) > 0.90   # LPT over whole sequences reaches 0.93 of the mean at best
            if world == 1:
                assert all(p["last"] and p["first_read"] == 0 for p in plan[0])
    # a sequence much longer than a rank's share is cut several times: all parts but the last are independent
    plan = SR.plan_line_split([1000000.0], 8)
    assert sum(len(p) for p in plan) == 8 and sum(1 for parts in plan for p in parts if p["last"]) == 1


def test_line_split_with_unequal_shares():
    """the host-delivery arm gives ranks behind a slower device-to-host path shorter pieces"""
    sizes = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64]
    reads = [s * 1000.0 for s in sizes]
    shares = [8.0, 8.0, 8.0, 8.0, 18.0, 18.0, 18.0, 18.0]
    plan = SR.plan_line_split(reads, 8, shares=shares)
    _check_plan(reads, 8, plan)
    loads = [sum(p["est"] for p in parts) for parts in plan]
    for ld, sh in zip(loads, shares):
        assert abs(ld / sum(loads) - sh / sum(shares)) < 0.02


def test_split_order_runs_feeders_first_and_dependents_last():
    plan = SR.plan_line_split([1000.0, 900.0, 800.0, 700.0], 3)
    for parts in plan:
        f, w, d = SR.split_order(parts)
        assert all(not p["last"] for p in f)
        assert all(p["last"] and p["first_read"] == 0 for p in w)
        assert all(p["last"] and p["first_read"] > 0 for p in d)
        assert len(f) + len(w) + len(d) == len(parts)


def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = SR.SplitExchange(3, dist)
    if rank == 0:
        ex.add(1, 12345)       # rank 0 simulated the first part of sequence 1 ...
        ex.add(2, 100)
    else:
        ex.add(2, 23)          # ... and both ranks a part of sequence 2
    ex.publish()
    q.put((rank, ex.prefix(0), ex.prefix(1), ex.prefix(2)))
    dist.destroy_process_group()


def test_split_exchange_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 0, 12345, 123), (1, 0, 12345, 123)]


class _FakeStats:
    def __init__(self, len_total_end, res_num):
        self.len_total_end, self.res_num = len_total_end, res_num
        self.kernel_launches = 1
        self.sim_seconds = self.emit_seconds = self.seg_seconds = self.chain_seconds = self.gen_seconds = 0.0


class _FakeWorkload:
    """stands in for bench.Workload: every read emits exactly 10 bases, a sequence's quota is 10 x its read count"""

    def __init__(self, reads_per_seq, rank):
        self.seqset, self.local, self.rank = None, 0, rank
        self.contigs = list(reads_per_seq)
        self.log = []

    def engine(self, lane):
        return lane

    def run_part(self, part, rng_seed=0, read_range=None, prefix=0, host_seq=None, sink=None, on_chunk=None, eng=None):
        import time
        time.sleep(0.01)
        total = self.contigs[part["seq"]]
        n = part["max_reads"] if not part["last"] else total - part["first_read"]
        # a dependent last part must have been told what the parts in front of it emitted
        assert prefix == (10 * part["first_read"] if part["last"] else 0), (part, prefix)
        self.log.append((part["seq"], part["first_read"], n, eng))
        return 10 * n, 0, _FakeStats(prefix + 10 * n if part["last"] else 10 * n, n)


def _run_parts_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    reads = [10000, 4000, 7000, 3000, 9000]
    plan = SR.plan_line_split([float(r) for r in reads], world)
    W = _FakeWorkload(reads, rank)
    got = []
    for lanes in (1, 2):
        bases = [0]
        bench.run_parts(W, plan[rank], dist, True, lambda p, b, ob, st: bases.__setitem__(0, bases[0] + b), lanes=lanes)
        got.append(bases[0])
    q.put((rank, got, sorted(W.log)))
    dist.destroy_process_group()


def test_bench_run_parts_orders_feeders_exchange_and_dependents_world2():
    """bench.run_parts on two ranks (gloo) with a fake workload, one and two lanes: every read of every sequence is
    simulated exactly once, and every dependent last part starts from the emitted bases of the parts in front of it"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_run_parts_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    reads = [10000, 4000, 7000, 3000, 9000]
    for lanes_i in (0, 1):
        assert sum(r[1][lanes_i] for r in res) == 10 * sum(reads)
    covered = {}
    for _, _, log in res:
        for seq, first, n, _ in log[::1]:
            covered.setdefault(seq, []).append((first, n))
    for k, total in enumerate(reads):
        spans = sorted(set(covered[k]))
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert any(len(set(v)) > 1 for v in covered.values())   # a sequence was shared by the two ranks


def test_line_split_plans_randomised():
    """random sequence counts, sizes (down to a handful of reads), rank counts and delivery shares: every read belongs
    to exactly one part, parts of a sequence are consecutive, only the last one runs to the quota"""
    import numpy as np
    rng = np.random.default_rng(2024)
    for _ in range(300):
        n = int(rng.integers(1, 30))
        world = int(rng.integers(1, 17))
        reads = [float(x) for x in np.exp(rng.uniform(np.log(3.0), np.log(3e5), n))]
        shares = None if rng.random() < 0.5 else [float(x) for x in rng.uniform(0.5, 3.0, world)]
        weights = None if rng.random() < 0.5 else [r * 50.0 + 4.5e4 for r in reads]
        plan = SR.plan_line_split(reads, world, weights=weights, shares=shares)
        assert len(plan) == world
        _check_plan(reads, world, plan)
        for parts in plan:
            for p in parts:
                assert p["first_read"] >= 0 and p["max_reads"] >= 0 and p["est"] >= 0.0
                assert (p["max_reads"] == 0) == p["last"]
