#!/usr/bin/env python
"""Host ceiling of the delivery path: N processes (one per GPU), each copying device memory into its own pinned
host buffers with back-to-back cudaMemcpyAsync (128 MiB pieces, two slots) for a few seconds, all at the same time.
Prints aggregate and per-GPU GB/s for N = 1, 2, 4, 8 (as many as there are GPUs).  Diagnostic, not the bench.

    python tools/d2h_ceiling.py [--seconds 3]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, n, seconds, pin, q, start_evt):
    torch.cuda.set_device(rank)
    if pin:
        import bench
        bench.pin_to_gpu_numa_node(rank)
    piece = 128 << 20
    dev = torch.empty(piece, dtype=torch.uint8, device="cuda")
    host = [torch.empty(piece, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for h in host:
            h.copy_(dev, non_blocking=True)
    st.synchronize()
    q.put(("ready", rank))
    start_evt.wait()
    t0 = time.perf_counter()
    nbytes = 0
    with torch.cuda.stream(st):
        while time.perf_counter() - t0 < seconds:
            for h in host:
                h.copy_(dev, non_blocking=True)
                nbytes += piece
            st.synchronize()
    dt = time.perf_counter() - t0
    q.put(("done", rank, nbytes / dt / 1e9))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    args = ap.parse_args()
    ngpu = torch.cuda.device_count()
    out = {"gpus_visible": ngpu, "cpus": os.cpu_count(), "runs": []}
    for pin in (False, True):
        for n in (1, 2, 4, 8):
            if n > ngpu:
                break
            ctx = mp.get_context("spawn")
            q = ctx.Queue()
            ev = ctx.Event()
            ps = [ctx.Process(target=worker, args=(r, n, args.seconds, pin, q, ev)) for r in range(n)]
            for p in ps:
                p.start()
            for _ in range(n):
                q.get()
            ev.set()
            rates = {}
            for _ in range(n):
                m = q.get()
                rates[m[1]] = m[2]
            for p in ps:
                p.join()
            out["runs"].append({"n": n, "numa_pinned": pin, "aggregate_gb_s": sum(rates.values()),
                                "per_gpu_gb_s": [round(rates[r], 2) for r in sorted(rates)]})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
