"""Multi-GPU plumbing of the path: reads shard across ranks with no data-path
collective; the ONE exchange is the statistics block (sim.res_* counters, freq_len, freq_accuracy —
pbsim.cpp:2293-2316, :2387-2410), all-reduced once per reference sequence, plus an all-gather of one
int64 per rank when a sequence's read-id range is split and the quota cut needs the global prefix.
torch.distributed is only the transport (NCCL on GPUs, gloo in the CPU tests)."""
import torch

# layout of the engine's stats block (emit.cuh k_stats): int64 cells
CELL_LEN_MIN = 3
CELL_LEN_MAX = 4


class _DeviceCells:
    """view of raw device memory for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr, cells):
        self.__cuda_array_interface__ = {"shape": (int(cells),), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2}


def reduce_stats_tensor(t, dist):
    """in-place: sum every cell over ranks, except res_len_min (MIN) and res_len_max (MAX)"""
    mn = t[CELL_LEN_MIN:CELL_LEN_MIN + 1].clone()
    mx = t[CELL_LEN_MAX:CELL_LEN_MAX + 1].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[CELL_LEN_MIN] = mn[0]
    t[CELL_LEN_MAX] = mx[0]
    return t


def allreduce_stats_block(engine, dist):
    ptr, cells = engine.stats_block()
    t = torch.as_tensor(_DeviceCells(ptr, cells), device="cuda")
    return reduce_stats_tensor(t, dist)


CELL_ACC_FX = 8          # sum of the per-read accuracies, 2^-40 fixed point (emit.cuh k_stats)
N_COUNTERS = 16


def merged_summary(cells, len_max, pass_num=1):
    """The reference's end-of-run numbers (pbsim.cpp:2387-2410, :5541-5564) from a REDUCED stats block (host array of
    int64 cells).  The accuracy mean comes from the fixed-point sum in the block: pbsim_stats.accuracy_total of one
    engine is the reference's floating-point sum in read order and cannot be combined across ranks."""
    import numpy as np
    c = np.asarray(cells, dtype=np.int64)
    res_num, res_pass, total = int(c[0]), int(c[1]), int(c[2])
    fa = c[N_COUNTERS:N_COUNTERS + 100001].astype(np.float64)
    fl = c[N_COUNTERS + 100001:].astype(np.float64)
    out = dict(res_num=res_num, res_pass_num=res_pass, res_len_total=total, res_len_min=int(c[CELL_LEN_MIN]),
               res_len_max=int(c[CELL_LEN_MAX]), res_sub_num=int(c[5]), res_ins_num=int(c[6]), res_del_num=int(c[7]))
    if res_pass > 0:
        out["res_len_mean"] = total / res_pass
        out["res_accuracy_mean"] = int(c[CELL_ACC_FX]) / 1099511627776.0 / res_pass
        if res_pass == 1:
            out["res_len_sd"] = out["res_accuracy_sd"] = 0.0
        else:
            i = np.arange(min(len(fl), len_max + 1), dtype=np.float64)
            out["res_len_sd"] = float(np.sqrt((((out["res_len_mean"] - i) ** 2) * fl[:len(i)]).sum() / res_pass))
            j = np.arange(100001, dtype=np.float64) * 0.00001
            out["res_accuracy_sd"] = float(np.sqrt((((out["res_accuracy_mean"] - j) ** 2) * fa).sum() / res_pass))
    return out


def lpt_assign(sizes, world):
    """by-sequence split of ONE run over the ranks: longest-processing-time-first, item indices per rank"""
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = loads.index(min(loads))
        loads[r] += sizes[i]
        out[r].append(i)
    return out


def contigs_for_rank(n_contigs, rank, world):
    """partition by sequence (many-contig genomes): rank r simulates contigs r, r+world, ..."""
    return list(range(rank, n_contigs, world))


def read_range_for_rank(n_reads, rank, world):
    """partition by read-index range inside one sequence: contiguous, sizes differ by at most 1"""
    base, extra = divmod(n_reads, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def emitted_prefix(my_emitted_bases, dist, device="cpu"):
    """exclusive scan over ranks of the emitted bases of each rank's read range: len_total_start of
    this rank, needed to place the quota cut (pbsim.cpp:2173-2181) when a sequence is split"""
    world = dist.get_world_size()
    mine = torch.tensor([int(my_emitted_bases)], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    vals = [int(v.item()) for v in allv]
    return sum(vals[:dist.get_rank()]), sum(vals)


# ---- one run split over the ranks inside sequences ("line split") ----------------------------------------------------
# The reference simulates a sequence until its quota of emitted bases is met (pbsim.cpp:2173-2181): how many reads a
# sequence gets is only known once every earlier read has been simulated.  Reads themselves depend on nothing but
# (seed, sequence, read number), so every part of a sequence EXCEPT ITS LAST can be simulated without knowing anything
# about the others (`max_reads` reads from `first_read`, the quota is far away); the last part runs to the quota and
# needs one number, the emitted bases of all parts before it (`len_total_start`).  The plan below lays all reads of the
# run on one line (sequences in order), cuts it into `world` pieces of equal estimated work, and lets every rank run
# its feeders first, the sequences it owns whole next and its dependent last parts at the end: by then the sums it
# needs have long been published by ONE asynchronous all-reduce of an int64 per sequence (SplitExchange).

def plan_line_split(reads_est, world, weights=None, snap_lo=0.02, snap_hi=0.04, shares=None, min_cut_reads=2000):
    """reads_est[k]: estimated read count of sequence k (quota / mean emitted bases per read); weights[k]: estimated
    work (default: reads_est).  Returns for every rank its parts in line order, each a dict
      seq, first_read, max_reads (0: run to the quota), last (the part that meets the quota), est (estimated work).
    A cut that falls into the first snap_lo or the last snap_hi of a sequence moves to the sequence's boundary: the
    estimate of a sequence's read count is good to a per cent or so, and a part that is not the last one must end
    safely in front of the quota; a sequence of fewer than min_cut_reads reads is never cut (the sum of a few hundred
    read lengths scatters by several per cent).  shares[r] (default: equal): the fraction of the work rank r should get — ranks
    whose delivery path is slower (GPUs behind a busier PCIe switch) get shorter pieces."""
    n = len(reads_est)
    w = [float(x) for x in (weights if weights is not None else reads_est)]
    total = sum(w)
    start = [0.0] * (n + 1)
    for k in range(n):
        start[k + 1] = start[k] + w[k]
    cut_in = {k: [] for k in range(n)}   # read indices at which sequence k is cut
    rank_cut_after = []  # for every rank boundary r (1..world-1): (k, read); (k, 0) = in front of sequence k
    if shares is None:
        shares = [1.0 / world] * world
    ssum = float(sum(shares))
    cum = 0.0
    for r in range(1, world):
        cum += shares[r - 1] / ssum
        x = total * cum
        k = 0
        while k + 1 < n and start[k + 1] <= x:
            k += 1
        f = (x - start[k]) / w[k] if w[k] > 0 else 0.0
        if reads_est[k] < min_cut_reads:
            f = 0.0 if f < 0.5 else 1.0
        if f < snap_lo:
            rank_cut_after.append((k, 0))
        elif f > 1.0 - snap_hi:
            rank_cut_after.append((k + 1, 0))
        else:
            rd = int(f * reads_est[k])
            if rd <= 0:
                rank_cut_after.append((k, 0))
            else:
                rank_cut_after.append((k, rd))
                if rd not in cut_in[k]:
                    cut_in[k].append(rd)
    out = [[] for _ in range(world)]
    for k in range(n):
        marks = [0] + sorted(cut_in[k])
        for i, lo in enumerate(marks):
            last = i + 1 == len(marks)
            hi = 0 if last else marks[i + 1]
            frac = ((reads_est[k] if last else hi) - lo) / max(1.0, float(reads_est[k]))
            piece = dict(seq=k, first_read=int(lo), max_reads=int(0 if last else hi - lo), last=last,
                         est=w[k] * max(0.0, frac))
            # owner: the number of rank boundaries at or in front of this piece's start
            r = sum(1 for (ck, crd) in rank_cut_after if (ck, crd) <= (k, lo))
            out[min(r, world - 1)].append(piece)
    return out


def split_order(parts):
    """execution order on one rank: feeders (parts others wait for), whole sequences, dependent last parts"""
    feeders = [p for p in parts if not p["last"]]
    whole = [p for p in parts if p["last"] and p["first_read"] == 0]
    dependent = [p for p in parts if p["last"] and p["first_read"] > 0]
    return feeders, whole, dependent


class SplitExchange:
    """the one data-dependent exchange of a line-split run: per sequence, the emitted bases of all parts that are not
    the last one.  Every rank adds what its feeders emitted and calls publish() once; the all-reduce runs
    asynchronously (NCCL on its own stream, gloo in the CPU tests) while the rank simulates the sequences it owns whole;
    prefix(k) waits for it (once) and returns len_total_start of sequence k's last part."""

    def __init__(self, n_seq, dist=None, device="cpu"):
        self.n, self.dist, self.device = n_seq, dist, device
        self.mine = [0] * n_seq
        self.work = self.t = self.sums = None

    def add(self, seq, emitted_bases):
        assert self.t is None, "SplitExchange.add after publish"
        self.mine[seq] += int(emitted_bases)

    def publish(self):
        self.t = torch.tensor(self.mine, dtype=torch.int64, device=self.device)
        if self.dist is not None:
            self.work = self.dist.all_reduce(self.t, op=self.dist.ReduceOp.SUM, async_op=True)

    def prefix(self, seq):
        if self.sums is None:
            if self.t is None:
                raise RuntimeError("SplitExchange.prefix before publish")
            if self.work is not None:
                self.work.wait()
            self.sums = [int(v) for v in self.t.cpu().tolist()]
        return self.sums[seq]
