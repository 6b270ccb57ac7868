"""Diagnostic (not the bench contract): throughput of --method sample on one GPU.

The pool is made the way a user would make it: reads simulated with QSHMM-RSII (mean 9 kb) are filtered by the host
front end (pbsim_host_sample_filter) and handed to the engine; then a synthetic 248 Mbp sequence is covered to
--depth 20 from that pool, records staying in HBM.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from pbsim_b200 import capi, simulator
    from tests.golden_util import model_path
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the engine has no CPU fallback")
    pool_mbp = int(os.environ.get("POOL_MBP", "30"))
    pool_depth = float(os.environ.get("POOL_DEPTH", "20"))
    L = capi.load()
    eng = simulator.Engine(0)
    eng.set_model(capi.HostModel(L, capi.host_params("qshmm"), model_path("QSHMM-RSII.model")))
    eng.set_synthetic_sequence(pool_mbp * 1000000, 1, 7)
    fastq, _, st, _ = eng.simulate(int(pool_depth * pool_mbp * 1000000), rng_mode=capi.RNG_PHILOX, seed=11)
    t0 = time.perf_counter()
    pool, ss = capi.sample_filter(L, fastq)
    t_filter = time.perf_counter() - t0
    hm = capi.HostModel(L, capi.host_params("sample"), None)
    eng.set_model(hm)
    eng.set_pool(pool)
    glen = 248000000
    out = []
    for step in range(3):
        eng.set_synthetic_sequence(glen, step + 1, 20240501 + step)
        torch.cuda.synchronize()
        eng.timer_start()
        eng.begin(int(20.0 * glen), rng_mode=capi.RNG_PHILOX, seed=1)
        bases = nbytes = 0
        while True:
            c = eng.next_chunk(device=True)
            if c is None:
                break
            bases += c.bases
            nbytes += c.reads_bytes + c.maf_bytes
        st = eng.end()
        ms = eng.timer_stop()
        out.append(dict(bases=bases, ms=ms, gbps=bases / ms / 1e6, sim_s=st.sim_seconds, emit_s=st.emit_seconds,
                        reads=st.res_num, launches=st.kernel_launches, out_bytes=nbytes))
    print(json.dumps({"diagnostic": "method sample, 248 Mbp synthetic sequence, depth 20, records in HBM",
                      "pool_reads": len(pool), "pool_bases": int(ss.len_total_filtered), "filter_seconds": t_filter,
                      "steps": out}))


if __name__ == "__main__":
    main()
