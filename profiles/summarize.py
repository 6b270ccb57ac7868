#!/usr/bin/env python
"""Reduce an .ncu-rep (ncu --set full) to the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== launch")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print("%-90s %s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE).stdout.decode()
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        c, s = Counter(), Counter()
        for r in rows[2:]:
            if len(r) <= max(ia, isrc, isamp) or not r[ia].isdigit():
                continue
            t = r[isrc].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            c[op] += int(r[ia])
            s[op] += int(r[isamp])
        tot, ts = sum(c.values()), max(1, sum(s.values()))
        print("== executed warp instructions by opcode (share of instructions / share of stall samples)")
        for op, n in c.most_common(14):
            print("%-8s %5.1f%% %5.1f%%" % (op, 100.0 * n / tot, 100.0 * s[op] / ts))


if __name__ == "__main__":
    main(sys.argv[1])
