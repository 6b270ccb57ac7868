#!/usr/bin/env python
"""bench.py — simulated Gbp/s of the PBSIM3 read-generation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4|c5] [--impl reference]

A STEP is one call of the reference's seam for one reference sequence: ingest the sequence
(get_genome_seq), then simulate_by_qshmm / simulate_by_errhmm to the depth quota, records emitted
(FASTQ + MAF).  Sequences are the 24 contigs of a synthetic 3.1 Gbp human-sized genome; step i works on
contig i mod 24, and with N GPUs every rank simulates its own range of read ids of that contig.  Default workload "c3" = BASELINE.json configs[2]
(WGS qshmm, QSHMM-ONT ultra-long reads, 3.1 Gbp genome, --depth 50): the configuration the metric
"simulated Gbp/s (WGS qshmm, 3.1 Gbp genome)" is quoted on.

  value : emitted bases / device time, sequence text already resident in HBM, records left in HBM
  e2e   : same steps through the C ABI with HOST buffers: the sequence text is uploaded from pinned
          host memory and every record byte is delivered to pinned host memory inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"

--impl reference times the UNMODIFIED reference binary (oracle/_ref/pbsim, compiled from
/root/reference/src/pbsim.cpp by `make -C oracle ref`) on the host cores, as N seed-split processes.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# GRCh38-like contig sizes (Mbp), total 3.088 Gbp, every sequence < 250 Mbp (reference limit 1e9, pbsim.cpp:24)
CONTIG_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47,
              51, 156, 57]
GENOME_SEED = 20240501

WORKLOADS = {
    # BASELINE.json configs[2]
    "c3": dict(name="WGS qshmm QSHMM-ONT ultra-long (mean 50 kb), 3.1 Gbp synthetic genome, --depth 50",
               method="qshmm", model="QSHMM-ONT.model", depth=50.0,
               params=dict(len_mean=50000.0, len_sd=35000.0, len_max=1000000, ratio=(39, 24, 36)),
               cli=["--length-mean", "50000", "--length-sd", "35000", "--length-max", "1000000",
                    "--difference-ratio", "39:24:36"]),
    # configs[0]'s model and defaults on the named genome size
    "c1": dict(name="WGS qshmm QSHMM-RSII defaults (mean 9 kb), 3.1 Gbp synthetic genome, --depth 20",
               method="qshmm", model="QSHMM-RSII.model", depth=20.0, params=dict(), cli=[]),
    # configs[1]
    "c2": dict(name="WGS errhmm ERRHMM-ONT-HQ (mean 9 kb), 3.1 Gbp synthetic genome, --depth 30",
               method="errhmm", model="ERRHMM-ONT-HQ.model", depth=30.0, params=dict(), cli=[]),
    # configs[4]: multi-pass CLR, SAM records (pbsim.cpp:2322-2333) instead of FASTQ
    "c5": dict(name="WGS errhmm ERRHMM-SEQUEL --pass-num 10 (mean 9 kb), 3.1 Gbp synthetic genome, --depth 20, SAM+MAF",
               method="errhmm", model="ERRHMM-SEQUEL.model", depth=20.0, params=dict(pass_num=10),
               cli=["--pass-num", "10"], batch_bases=3 << 30),
    # configs[3]: transcriptome; a step is one run over the whole transcript table
    "c4": dict(name="trans qshmm QSHMM-RSII, 200,000 synthetic transcripts (log-normal, median 1.5 kb), Zipf "
                    "expression, 2e7 reads",
               method="qshmm", model="QSHMM-RSII.model", depth=0.0, params=dict(), cli=[], strategy="trans",
               n_transcripts=200000, n_reads=20000000),
}

# DRAM traffic rate (dram__bytes_read.sum + dram__bytes_write.sum over gpu__time_duration.sum, GB/s) of the two big
# kernels from the committed ncu captures profiles/r01_k_sim_seg_c3_v17.txt / r01_k_emit_c3_v17.txt
NCU_DRAM_GBS = {("qshmm", "seg"): 567.0, ("qshmm", "emit"): 1602.6,
                # profiles/r01_k_sim_seg_err_c2_v13.txt / r01_k_emit_c2_v13.txt
                ("errhmm", "seg"): 366.7, ("errhmm", "emit"): 1385.7}

# algorithmic bytes per emitted base: FASTQ 2.002 + MAF 2.122 written + 0.244 read (2-bit genome)
ALGO_BYTES_PER_BASE = 4.37


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (recipe in B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.path = os.path.join(tempfile.gettempdir(), "pbsim_clocks_%d_%d.csv" % (os.getpid(), gpu_index))
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 9:
                        continue
                    try:
                        sm.append(float(p[1]))
                        smax.append(float(p[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                       p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on the host cores
# ------------------------------------------------------------------------------------------------
def _ref_sample_genome(path, mbp=5):
    import numpy as np
    rng = np.random.default_rng(GENOME_SEED)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=mbp * 1000000)].tobytes()
    with open(path, "wb") as f:
        f.write(b">sample\n")
        for i in range(0, len(s), 70):
            f.write(s[i:i + 70] + b"\n")
    return len(s)


def _parse_ref_bases(stderr):
    total = 0
    num = None
    for line in stderr.splitlines():
        if line.startswith("read num. :"):
            num = int(line.split(":")[1])
        elif line.startswith("read length mean (SD) :") and num is not None:
            total += int(round(num * float(line.split(":")[1].split("(")[0])))
            num = None
    return total


def run_reference_processes(wl, nproc, depth, workdir, seed0=1):
    """Launch nproc seed-split reference processes (--seed s+i) on the sample genome; returns (bases, wall)."""
    from oracle import refrun as R
    from tests.golden_util import model_path
    fa = os.path.join(workdir, "sample.fa")
    if not os.path.exists(fa):
        _ref_sample_genome(fa)
    env = dict(os.environ)
    env["PATH"] = R.SHIMS + ":" + env.get("PATH", "")  # gzip -> cat: generation + text formatting only
    args = ["--strategy", "wgs", "--method", wl["method"], "--" + wl["method"], model_path(wl["model"]),
            "--genome", fa, "--depth", str(depth)] + wl["cli"]
    procs = []
    t0 = time.perf_counter()
    for i in range(nproc):
        d = os.path.join(workdir, "p%d" % i)
        os.makedirs(d, exist_ok=True)
        procs.append(subprocess.Popen([R.REF_BIN] + args + ["--seed", str(seed0 + i), "--prefix", "o"], cwd=d, env=env,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    bases = 0
    for p in procs:
        _, err = p.communicate()
        bases += _parse_ref_bases(err.decode(errors="replace"))
    wall = time.perf_counter() - t0
    for i in range(nproc):
        d = os.path.join(workdir, "p%d" % i)
        for fn in os.listdir(d):
            try:
                os.remove(os.path.join(d, fn))
            except OSError:
                pass
    return bases, wall


def cpu_baseline(wl, kind_wanted="reference"):
    """Single-process reference on a bounded sample (about 10-20 s of CPU work)."""
    from oracle import refrun as R
    work = tempfile.mkdtemp(prefix="pbsim_cpu_")
    if R.have_reference_binary():
        depth = 24
        bases, wall = run_reference_processes(wl, 1, depth, work)
        return {"value": bases / wall / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "reference",
                "sample": "unmodified reference binary (g++ -O2), 1 process, 5 Mbp synthetic contig, --depth %d "
                          "(%d bases in %.1f s), gzip children replaced by cat (generation + formatting only)"
                          % (depth, bases, wall)}
    # the reference binary did not travel: time the C restatement instead
    from oracle import oracle as O
    from tests.golden_util import model_path
    import numpy as np
    o = O.Oracle(wl["method"], model_path(wl["model"]), **wl["params"])
    o.rng_glibc(1)
    rng = np.random.default_rng(GENOME_SEED)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5000000)].tobytes()
    o.set_sequence(s, 1)
    t0 = time.perf_counter()
    _, _, st = o.simulate_wgs(12)
    wall = time.perf_counter() - t0
    return {"value": st.res_len_total / wall / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port",
            "sample": "oracle C restatement, 5 Mbp synthetic contig, --depth 12 (%d bases in %.1f s)"
                      % (st.res_len_total, wall)}


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refrun as R
    base = {"impl": "reference", "metric": "simulated Gbp/s", "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": {"workload": wl["name"]}}
    if not R.have_reference_binary():
        cb = cpu_baseline(wl)
        base.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(base))
        return
    nproc = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="pbsim_refarm_")
    depth = 3  # per process and step: 5 Mbp x 3 = 15 Mbase, a couple of seconds
    for w in range(args.warmup):
        run_reference_processes(wl, nproc, depth, work, seed0=1000 + w * nproc)
    tot_b, tot_t = 0, 0.0
    for k in range(args.steps):
        b, t = run_reference_processes(wl, nproc, depth, work, seed0=1 + k * nproc)
        tot_b += b
        tot_t += t
    v = tot_b / tot_t / 1e9
    base.update(value=v, ms_per_step=tot_t / max(1, args.steps) * 1e3,
                cpu_baseline={"value": v, "unit": "Gbp/s", "cores": nproc, "kind": "reference",
                              "sample": "unmodified reference binary, %d seed-split processes per step, each 5 Mbp "
                                        "synthetic contig --depth %d; gzip children replaced by cat" % (nproc, depth)},
                e2e={"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    base["config"].update(sample_genome_bp=5000000, processes=nproc)
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink contigs (debugging only; reported in config)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the host-buffer arm (default: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--len-sd", type=float, default=None, help="diagnostics: override --length-sd (0 = equal-length reads)")
    ap.add_argument("--len-mean", type=float, default=None, help="diagnostics: override --length-mean")
    ap.add_argument("--batch-bases", type=float, default=None, help="diagnostics: engine target_batch_bases")
    ap.add_argument("--chain-chunk", type=int, default=None, help="diagnostics: engine chain_chunk (segments per chunk)")
    ap.add_argument("--first-batch-div", type=int, default=None, help="diagnostics: engine first_batch_div")
    ap.add_argument("--host-batch-bases", type=float, default=None, help="diagnostics: engine host_batch_bases")
    ap.add_argument("--bam", action="store_true", help="multi-pass workloads: BAM records / BGZF blocks instead of SAM text")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.len_sd is not None or args.len_mean is not None:
        wl["params"] = dict(wl["params"])
        if args.len_sd is not None:
            wl["params"]["len_sd"] = args.len_sd
        if args.len_mean is not None:
            wl["params"]["len_mean"] = args.len_mean
        wl["name"] += " [diagnostic override: len_mean=%s len_sd=%s]" % (args.len_mean, args.len_sd)
    if args.impl == "reference":
        reference_arm(args, wl)
        return

    import numpy as np
    import torch
    from pbsim_b200 import capi, simulator
    from tests.golden_util import model_path

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = capi.load()
    hm = capi.HostModel(L, capi.host_params(wl["method"], **wl["params"]), model_path(wl["model"]))
    eng = simulator.Engine(local)
    eng.set_model(hm)
    if args.batch_bases:
        eng.set_option("target_batch_bases", int(args.batch_bases))
    if args.chain_chunk:
        eng.set_option("chain_chunk", args.chain_chunk)
    if args.first_batch_div:
        eng.set_option("first_batch_div", args.first_batch_div)
    if args.host_batch_bases:
        eng.set_option("host_batch_bases", int(args.host_batch_bases))
    if args.bam:
        eng.set_option("bam", 1)
        wl["name"] += " [BAM records]"
    if wl.get("batch_bases") and not args.batch_bases:
        eng.set_option("target_batch_bases", int(wl["batch_bases"]))
    depth = wl["depth"]
    contigs = [max(200000, int(m * 1000000 * args.scale)) for m in CONTIG_MBP]
    bias = [0.0] + [1.0] * 10 + [0.0]
    seqset = None
    if wl.get("strategy") == "trans":
        # log-normal lengths (median 1.5 kb, capped), Zipf expression counts, both strands
        rng = np.random.default_rng(GENOME_SEED + rank)
        nt = max(100, int(wl["n_transcripts"] * args.scale))
        lens = np.clip(np.exp(rng.normal(np.log(1500.0), 0.75, nt)).astype(np.int64), 200, 100000)
        w = 1.0 / np.arange(1, nt + 1) ** 0.9
        expr = rng.permutation(np.maximum(1, (w / w.sum() * wl["n_reads"] * args.scale)).astype(np.int64))
        text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(lens.sum()))].tobytes()
        st0 = np.zeros(nt + 1, dtype=np.int64)
        st0[1:] = np.cumsum(lens)
        plus = (expr + 1) // 2
        names = [b"T%06d" % (t + 1) for t in range(nt)]
        id_start = np.zeros(nt + 1, dtype=np.int32)
        id_start[1:] = np.cumsum([len(x) for x in names])
        seqset = dict(n=nt, bases=text, start=st0, plus=plus.astype(np.int32), minus=(expr - plus).astype(np.int32),
                      ids=b"".join(names), id_start=id_start)
        contigs = [int(lens.sum())]
        eng.set_seqset("trans", seqset, bias)

    # multi-GPU: reads shard by read-id range with no data-path collective.  Every rank walks
    # the same contigs and simulates its own range of read ids of each (first_read = rank << 26; results depend only
    # on (seed, sequence, read id)), to the full depth quota: per-GPU work is fixed as N grows (weak scaling) and
    # equal across ranks.
    first_read = rank << 26

    def seq_of(step):
        return step % len(contigs)

    def step_device(step):
        """ingest (synthetic text generated in HBM) + simulate to the quota, records stay in HBM"""
        k = seq_of(step)
        if seqset is None:
            eng.set_synthetic_sequence(contigs[k], k + 1, GENOME_SEED + k)
        eng.begin(int(depth * contigs[k]), rng_mode=capi.RNG_PHILOX, seed=1 + ((step + 1000 * rank) if seqset else 0),
                  first_read=first_read if seqset is None else 0)
        bases = out_bytes = 0
        while True:
            c = eng.next_chunk(device=True)
            if c is None:
                break
            bases += c.bases
            out_bytes += c.reads_bytes + c.maf_bytes
        st = eng.end()
        return bases, out_bytes, st

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows every arena to its steady-state size)
    for w in range(args.warmup):
        step_device(w)
    # ---- timed: device-resident arm
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    eng.timer_start()
    t0 = time.perf_counter()
    bases = out_bytes = launches = 0
    sim_s = emit_s = gen_s = seg_s = chain_s = 0.0
    for k in range(args.steps):
        b, ob, st = step_device(args.warmup + k)
        bases += b
        out_bytes += ob
        launches += st.kernel_launches
        sim_s += st.sim_seconds
        emit_s += st.emit_seconds
        seg_s += st.seg_seconds
        chain_s += st.chain_seconds
        gen_s += st.gen_seconds
    dev_ms = eng.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()

    # ---- timed: end-to-end arm (host buffers both ways)
    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    e2e = None
    if e2e_steps > 0:
        need = sorted({seq_of(args.warmup + k) for k in range(e2e_steps)}) if seqset is None else []
        host_seq = {}
        for k in need:  # pinned host copies of the sequence text (untimed preparation)
            eng.set_synthetic_sequence(contigs[k], k + 1, GENOME_SEED + k)
            t = torch.empty(contigs[k], dtype=torch.uint8, pin_memory=True)
            eng.get_sequence_ascii(t.data_ptr(), contigs[k])
            host_seq[k] = t
        def e2e_pass():
            barrier()
            t0 = time.perf_counter()
            e_bases = h2d = d2h = 0
            sink = 0
            for k in range(e2e_steps):
                q = seq_of(args.warmup + k)
                if seqset is None:
                    eng.set_sequence_ptr(host_seq[q].data_ptr(), contigs[q], q + 1, bias)
                else:
                    eng.set_seqset("trans", seqset, bias)  # the table's text goes up with every step
                h2d += contigs[q]
                eng.begin(int(depth * contigs[q]), rng_mode=capi.RNG_PHILOX, seed=1 + (rank if seqset else 0),
                          first_read=first_read if seqset is None else 0)
                while True:
                    c = eng.next_chunk(device=False)
                    if c is None:
                        break
                    e_bases += c.bases
                    d2h += c.reads_bytes + c.maf_bytes
                    if c.reads_bytes:  # the consumer looks at the delivered bytes
                        sink ^= C.c_ubyte.from_address(c.reads + c.reads_bytes - 1).value
                eng.end()
            barrier()
            e_ms = (time.perf_counter() - t0) * 1e3
            return dict(bases=e_bases, ms=e_ms, h2d=h2d / e2e_steps, d2h=d2h / e2e_steps)

        e2e = e2e_pass()
        # the same with the records gzip-compressed on the GPU before they cross PCIe (the reference's outputs are
        # .gz files, pbsim.cpp:708-730); reported next to the text number, not instead of it
        eng.set_option("deflate", 1)
        e2e_gz = e2e_pass()
        eng.set_option("deflate", 0)
        del host_seq

    # ---- reduce over ranks: total units, MAX time
    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    tot_bases = allsum(bases)
    max_ms = allmax(dev_ms)
    max_wall = allmax(wall_ms)
    tot_launches = allsum(launches)
    if e2e:
        e2e_bases = allsum(e2e["bases"])
        e2e_ms = allmax(e2e["ms"])
        e2e_gz_bases = allsum(e2e_gz["bases"])
        e2e_gz_ms = allmax(e2e_gz["ms"])
    if dist is not None:
        # the one collective of the path: statistics block (counters + both histograms), NCCL all-reduce
        from pbsim_b200.stats_reduce import allreduce_stats_block
        allreduce_stats_block(eng, dist)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = tot_bases / (max_ms * 1e-3) / 1e9
        kern_s = sim_s + emit_s
        # the dominant kernel is pass 2 (k_emit): it performs the path's algorithmic traffic (genome read, records
        # written); its launches are bracketed by CUDA events on the engine's stream (pbsim_stats.emit_seconds)
        seg_name = "k_sim_seg" if wl["method"] == "qshmm" else "k_sim_seg_err"
        dom_s, dom_name = (seg_s, seg_name + " (pass 1, segment-parallel chains)") if seg_s >= emit_s else \
                          (emit_s, "k_emit (pass 2)")
        achieved = ALGO_BYTES_PER_BASE * bases / dom_s / 1e9 if dom_s > 0 else 0.0
        path_achieved = ALGO_BYTES_PER_BASE * bases / kern_s / 1e9 if kern_s > 0 else 0.0
        traffic = NCU_DRAM_GBS.get((wl["method"], "seg" if seg_s >= emit_s else "emit"))
        line = {
            "metric": "simulated Gbp/s", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl["name"],
                       "step": ("one run over the transcript table, FASTQ+MAF emitted" if seqset is not None else
                                "one reference sequence: device ingest + simulate to depth quota, %s+MAF emitted"
                                % ("SAM" if wl["params"].get("pass_num", 1) > 1 else "FASTQ")), "genome_bp": int(sum(contigs)), "contigs": len(contigs), "rng": "philox4x32-10",
                       "l2": "every step writes > 10 GB of records and events (>> 126 MB L2); no explicit flush needed",
                       "host_wall_ms_per_step": max_wall / max(1, args.steps), "scale": args.scale,
                       "sharding": "every rank simulates its own read-id range (first_read = rank << 26) of the same "
                                   "contig to the full depth quota; no data-path collective, one NCCL all-reduce of "
                                   "the statistics block"},
            "gpu_launches": int(tot_launches),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of that "
                                           "kernel / its duration, GB/s (ncu --set full, profiles/r01_k_*_c3_v17.txt, *_c2_v13.txt)",
                         "kernel": dom_name + ", rank 0",
                         "algorithmic_bytes_per_base": ALGO_BYTES_PER_BASE, "peak_source": peak_src,
                         "kernel_seconds": {"sim": sim_s, "of_which_" + seg_name: seg_s,
                                            "of_which_k_chain_chunk": chain_s, "emit": emit_s,
                                            "all_generation": gen_s},
                         "kernel_share_of_step": dom_s / (dev_ms * 1e-3) if dev_ms else None,
                         "path": {"what": "pass 1 + pass 2 together (k_sim_seg, k_find_end, k_sim_%s, k_emit)"
                                          % wl["method"],
                                  "achieved": path_achieved, "frac": path_achieved / peak if peak else None,
                                  "share_of_step": kern_s / (dev_ms * 1e-3) if dev_ms else None}},
        }
        if e2e:
            line["e2e"] = {"value": e2e_bases / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "delivered": "text records"}
            line["e2e_gzip"] = {"value": e2e_gz_bases / (e2e_gz_ms * 1e-3) / 1e9, "unit": "Gbp/s",
                                "h2d_bytes_per_step": e2e_gz["h2d"], "d2h_bytes_per_step": e2e_gz["d2h"],
                                "steps": e2e_steps, "ms_per_step": e2e_gz_ms / e2e_steps,
                                "delivered": "gzip members written by the GPU (option deflate)",
                                "compression_ratio": e2e["d2h"] / max(1.0, e2e_gz["d2h"])}
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(wl)
            except Exception as ex:  # never lose the GPU line to a host-side hiccup
                line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": 1, "kind": "reference",
                                        "sample": "failed: %r" % (ex,)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
