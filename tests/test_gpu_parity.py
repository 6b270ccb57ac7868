"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libpbsim_cuda.so); the oracle / golden fixtures are only the checker.

  replay mode  : engine fed the reference's own rand() draws  == reference bytes (tests/golden)
  philox mode  : engine                                        == CPU oracle running the same Philox addressing
"""
import numpy as np
import pytest

from oracle import oracle as O
from pbsim_b200 import capi, simulator
from tests.golden_util import Case, case_names
from tests.gpu_util import engine_model, run_case_on_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = simulator.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("name", case_names())
def test_replay_reproduces_reference_bytes(eng, name):
    c = Case(name)
    out, _ = c.run_oracle("glibc")
    res = run_case_on_gpu(c, eng, "replay", oracle_out=out)
    for i, (reads, maf, st, text) in enumerate(res, start=1):
        assert reads == c.reads(i), "FASTQ/SAM differs from the reference, seq %d" % i
        assert maf == c.maf(i), "MAF differs from the reference, seq %d" % i
        assert text == c.stats_blocks[i], "stats block differs, seq %d" % i


@pytest.mark.parametrize("name", case_names())
def test_philox_equals_oracle_philox(eng, name):
    c = Case(name)
    out, _ = c.run_oracle("philox")
    res = run_case_on_gpu(c, eng, "philox")
    for i, ((reads, maf, st, text), o) in enumerate(zip(res, out), start=1):
        assert reads == o["reads"], "reads differ from the oracle, seq %d" % i
        assert maf == o["maf"], "maf differs from the oracle, seq %d" % i
        assert text == o["stats_text"]


@pytest.mark.parametrize("name", ["qs_rsii_quirks", "err_sequel_multipass"])
@pytest.mark.parametrize("batch", [1, 7, 64])
def test_output_is_independent_of_batching(eng, name, batch):
    """chunking (and therefore the speculative-plan / quota-cut / tail logic) must not change a byte"""
    c = Case(name)
    ref = run_case_on_gpu(c, eng, "philox")
    got = run_case_on_gpu(c, eng, "philox", batch_reads=batch)
    for a, b in zip(ref, got):
        assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3]


def test_replay_batched_still_reproduces_reference(eng):
    c = Case("qs_rsii_basic")
    out, _ = c.run_oracle("glibc")
    res = run_case_on_gpu(c, eng, "replay", oracle_out=out, batch_reads=13)
    for i, (reads, maf, st, text) in enumerate(res, start=1):
        assert reads == c.reads(i) and maf == c.maf(i) and text == c.stats_blocks[i]


def test_histograms_equal_oracle(eng):
    c = Case("qs_rsii_basic")
    out, _ = c.run_oracle("philox")
    hm = engine_model(c)
    eng.set_model(hm)
    _, s = c.contigs[0]
    eng.set_sequence(s, 1, [0.0] + [1.0] * 10 + [0.0])
    reads, maf, (st, fl, fa), n = eng.simulate(int(c.depth * len(s)), rng_mode=capi.RNG_PHILOX, seed=c.seed,
                                               want_hist=True)
    assert np.array_equal(fl[:len(out[0]["freq_len"])], out[0]["freq_len"][:len(fl)])
    assert np.array_equal(fa, out[0]["freq_accuracy"])
    assert st.res_num == out[0]["stats"].res_num
    assert st.res_sub_num == out[0]["stats"].res_sub_num
    assert st.accuracy_total == out[0]["stats"].accuracy_total


def test_genome_ingest_matches_oracle_hp(eng):
    """K0: hpfreq of the device ingest == the oracle's get_genome_seq restatement"""
    c = Case("qs_rsii_quirks")
    o = c.new_oracle()
    hm = engine_model(c)
    eng.set_model(hm)
    for i, (_, s) in enumerate(c.contigs, start=1):
        o.set_sequence(s, i)
        hp = o.hp()
        want = [int(np.count_nonzero(hp == h)) for h in range(12)]
        eng.set_sequence(s, i, [0.0] + [1.0] * 10 + [0.0])
        assert eng.hpfreq() == want


def test_host_delivery_in_small_pieces_is_byte_identical(eng):
    """pinned staging of 4 KiB: the records arrive as many pieces; concatenation must not change a byte"""
    c = Case("qs_rsii_basic")
    ref = run_case_on_gpu(c, eng, "philox")
    eng.set_option("stage_bytes", 4096)
    try:
        got = run_case_on_gpu(c, eng, "philox")
    finally:
        eng.set_option("stage_bytes", 128 << 20)
    for a, b in zip(ref, got):
        assert a[0] == b[0] and a[1] == b[1]


def test_synthetic_sequence_roundtrip_properties(eng):
    """size-independent properties on a device-generated 3 Mbp contig: every MAF block re-derives the FASTQ
    read (read row minus gaps == read, reverse-complemented for '-') and the genome window (ref row minus gaps)"""
    import ctypes as C
    c = Case("qs_rsii_basic")
    hm = engine_model(c)
    eng.set_model(hm)
    n = 3000000
    eng.set_synthetic_sequence(n, 1, 99)
    buf = (C.c_char * n)()
    eng.get_sequence_ascii(C.addressof(buf), n)
    genome = bytes(buf)
    reads, maf, st, _ = eng.simulate(2 * n, rng_mode=capi.RNG_PHILOX, seed=5)
    fq = reads.split(b"\n")
    blocks = maf.split(b"\n\n")[:-1]
    assert len(fq) // 4 == len(blocks) == st.res_num
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    total = 0
    for k, blk in enumerate(blocks):
        l1, l2 = blk.split(b"\n")[1:3]
        f1, f2 = l1.split(), l2.split()
        off, wlen, refrow = int(f1[2]), int(f1[3]), f1[6]
        rlen, strand, readrow = int(f2[3]), f2[4], f2[6]
        assert len(refrow) == len(readrow)
        assert refrow.replace(b"-", b"") == genome[off:off + wlen]
        seq = fq[4 * k + 1]
        assert len(seq) == rlen == len(fq[4 * k + 3])
        rr = readrow.replace(b"-", b"")
        assert (rr if strand == b"+" else rr.translate(comp)[::-1]) == seq
        total += rlen
    assert total == st.res_len_total >= 2 * n


@pytest.mark.parametrize("name", ["qs_rsii_basic", "qs_rsii_quirks", "qs_delheavy_uniform", "qs_rsii_multipass",
                                  "qs_ont_hpbias", "qs_hp11_uniform", "err_onthq_basic", "err_sequel_multipass",
                                  "err_sequel_hiacc"])
def test_segment_parallel_pass1_equals_oracle(eng, name):
    """segment-parallel pass 1 forced onto every read longer than one segment: same bytes as the oracle"""
    c = Case(name)
    out, _ = c.run_oracle("philox")
    eng.set_option("seg_min_len", 1024)
    try:
        res = run_case_on_gpu(c, eng, "philox")
    finally:
        eng.set_option("seg_min_len", 2048)
    for i, ((reads, maf, st, text), o) in enumerate(zip(res, out), start=1):
        assert reads == o["reads"], "reads differ from the oracle, seq %d" % i
        assert maf == o["maf"], "maf differs from the oracle, seq %d" % i
        assert text == o["stats_text"]


@pytest.mark.parametrize("method,model", [("qshmm", "QSHMM-RSII.model"), ("errhmm", "ERRHMM-ONT-HQ.model")])
def test_segments_on_and_off_give_identical_bytes(eng, method, model):
    """3 Mbp device-generated contig, default lengths (mean 9 kb): the segment-parallel and the sequential pass 1
    must agree byte for byte, statistics included"""
    from tests.golden_util import model_path
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params(method), model_path(model))
    eng.set_model(hm)
    n = 3000000
    eng.set_synthetic_sequence(n, 1, 7)
    outs = []
    for on in (1, 0):
        eng.set_option("segments", on)
        reads, maf, (st, fl, fa), _ = eng.simulate(3 * n, rng_mode=capi.RNG_PHILOX, seed=11, want_hist=True)
        outs.append((reads, maf, st.res_num, st.res_len_total, st.accuracy_total, st.res_sub_num, st.res_ins_num,
                     st.res_del_num, fa.tobytes()))
    eng.set_option("segments", 1)
    assert outs[0][2] > 500
    assert outs[0] == outs[1]


def _device_bytes(ptr, n):
    """copy n bytes out of HBM with the CUDA runtime (device delivery hands out raw device pointers)"""
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    buf = C.create_string_buffer(n)
    assert rt.cudaMemcpy(buf, ptr, n, 2) == 0
    return buf.raw[:n]


@pytest.mark.parametrize("method,model", [("qshmm", "QSHMM-RSII.model"), ("errhmm", "ERRHMM-ONT-HQ.model")])
def test_pipelined_delivery_is_byte_identical(eng, method, model):
    """the producer thread (batch k+1 generated while batch k leaves HBM) must not change a byte: host delivery
    with the pipeline off / on, many small batches and pieces, and pipelined device delivery"""
    from tests.golden_util import model_path
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params(method), model_path(model))
    eng.set_model(hm)
    n = 2000000
    eng.set_synthetic_sequence(n, 1, 13)
    outs = []
    try:
        eng.set_option("stage_bytes", 1 << 16)
        for pipe in (0, 1):
            eng.set_option("pipeline", pipe)
            reads, maf, st, nchunks = eng.simulate(3 * n, rng_mode=capi.RNG_PHILOX, seed=3, batch_reads=64)
            outs.append((reads, maf, st.res_num, st.res_len_total, st.accuracy_total))
            assert nchunks > 20
        # device delivery, pipelined: the chunk of batch k stays valid while batch k+1 is being generated
        eng.set_option("pipeline", 2)
        eng.begin(3 * n, rng_mode=capi.RNG_PHILOX, seed=3, batch_reads=64)
        reads, maf, nb = [], [], 0
        while True:
            c = eng.next_chunk(device=True)
            if c is None:
                break
            assert c.on_device == 1
            reads.append(_device_bytes(c.reads, c.reads_bytes))
            maf.append(_device_bytes(c.maf, c.maf_bytes))
            nb += 1
        st = eng.end()
        assert nb > 5
        outs.append((b"".join(reads), b"".join(maf), st.res_num, st.res_len_total, st.accuracy_total))
    finally:
        eng.set_option("pipeline", 1)
        eng.set_option("stage_bytes", 128 << 20)
    assert outs[0][2] > 300
    assert outs[0] == outs[1] == outs[2]


def test_abandoned_run_then_new_run(eng):
    """begin / a few chunks / begin again without draining: the producer is stopped and the next run is clean"""
    c = Case("qs_rsii_basic")
    ref = run_case_on_gpu(c, eng, "philox")
    hm = engine_model(c)
    eng.set_model(hm)
    eng.set_sequence(c.contigs[0][1], 1, [0.0] + [1.0] * 10 + [0.0])
    eng.begin(10 * len(c.contigs[0][1]), rng_mode=capi.RNG_PHILOX, seed=c.seed, batch_reads=8)
    assert eng.next_chunk() is not None
    eng.end()
    eng.begin(10 * len(c.contigs[0][1]), rng_mode=capi.RNG_PHILOX, seed=c.seed, batch_reads=8)
    assert eng.next_chunk() is not None
    got = run_case_on_gpu(c, eng, "philox")  # begins again over the abandoned run
    for a, b in zip(ref, got):
        assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3]


def test_read_range_shards_concatenate_to_the_single_gpu_run(eng):
    """multi-GPU sharding of one sequence by read-id range: rank 0 takes reads [0, k), rank 1
    the rest, told how many bases rank 0 emitted (the all_gather of stats_reduce.emitted_prefix); the quota cut and
    every byte must equal the one-engine run"""
    from tests.golden_util import model_path
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params("qshmm"), model_path("QSHMM-RSII.model"))
    eng.set_model(hm)
    n = 1500000
    eng.set_synthetic_sequence(n, 1, 21)
    quota = 3 * n
    whole = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=4)
    k = whole[2].res_num // 3
    a = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=4, first_read=0, max_reads=k)
    b = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=4, first_read=k, len_total_start=a[2].res_len_total)
    assert a[2].res_num == k and a[2].res_num + b[2].res_num == whole[2].res_num
    assert a[0] + b[0] == whole[0]
    assert a[1] + b[1] == whole[1]
    assert a[2].res_len_total + b[2].res_len_total == whole[2].res_len_total


def test_merged_statistics_of_a_split_run_equal_the_single_engine_statistics(eng):
    """the stats blocks of a 2-way read-range split, summed as the NCCL all-reduce sums them, give the one-engine
    statistics text: the accuracy mean comes from the block's fixed-point sum (pbsim_stats.accuracy_total, the
    reference's floating-point sum in read order, cannot be combined across ranks)"""
    import ctypes as C
    import torch
    from pbsim_b200 import stats_reduce as SR
    from tests.golden_util import model_path
    hm = capi.HostModel(capi.load(), capi.host_params("qshmm"), model_path("QSHMM-RSII.model"))
    eng.set_model(hm)
    n = 1200000
    eng.set_synthetic_sequence(n, 1, 33)
    quota = 3 * n

    def block():
        ptr, cells = eng.stats_block()
        return torch.as_tensor(SR._DeviceCells(ptr, cells), device="cuda").cpu().numpy().copy()

    whole = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=9)
    blk_whole = block()
    k = whole[2].res_num // 2
    a = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=9, first_read=0, max_reads=k)
    blk_a = block()
    b = eng.simulate(quota, rng_mode=capi.RNG_PHILOX, seed=9, first_read=k, len_total_start=a[2].res_len_total)
    blk_b = block()
    merged = blk_a + blk_b
    merged[SR.CELL_LEN_MIN] = min(blk_a[SR.CELL_LEN_MIN], blk_b[SR.CELL_LEN_MIN])
    merged[SR.CELL_LEN_MAX] = max(blk_a[SR.CELL_LEN_MAX], blk_b[SR.CELL_LEN_MAX])
    assert (merged == blk_whole).all()          # every cell, the fixed-point accuracy sum included
    m = SR.merged_summary(merged, hm.view.len_max)
    st = whole[2]
    assert m["res_num"] == st.res_num and m["res_len_total"] == st.res_len_total
    assert abs(m["res_accuracy_mean"] - st.res_accuracy_mean) < 1e-9
    assert abs(m["res_accuracy_sd"] - st.res_accuracy_sd) < 1e-9
    assert abs(m["res_len_sd"] - st.res_len_sd) < 1e-6 * max(1.0, st.res_len_sd)
    assert "%f (%f)" % (m["res_accuracy_mean"], m["res_accuracy_sd"]) == "%f (%f)" % (st.res_accuracy_mean, st.res_accuracy_sd)


def test_line_split_of_a_multi_sequence_run_concatenates_to_the_single_gpu_bytes(eng):
    """bench.py --gpus N / stats_reduce.plan_line_split: ONE run of several sequences cut into `world` pieces of equal
    estimated work.  The ranks are emulated one after the other on this GPU in the order a real run imposes (every
    rank's feeders, the exchange of emitted bases, then whole sequences and dependent last parts); the parts of every
    sequence, concatenated in read order, must be the bytes of the sequence simulated whole, and the summed statistics
    blocks the whole run's."""
    from pbsim_b200 import stats_reduce as SR
    from tests.golden_util import model_path
    hm = capi.HostModel(capi.load(), capi.host_params("qshmm"), model_path("QSHMM-RSII.model"))
    eng.set_model(hm)
    lens = [900000, 400000, 1300000, 700000, 250000]
    depth = 3

    def run(k, **kw):
        eng.set_synthetic_sequence(lens[k], k + 1, 100 + k)
        return eng.simulate(depth * lens[k], rng_mode=capi.RNG_PHILOX, seed=6, **kw)

    whole = [run(k) for k in range(len(lens))]
    pilot = run(0, max_reads=64)
    mean_emit = pilot[2].len_total_end / pilot[2].res_num
    for world in (2, 3, 7):
        # (sequences of a few hundred reads: the planner would not cut them on its own)
        plan = SR.plan_line_split([depth * n / mean_emit for n in lens], world, weights=[float(n) for n in lens],
                                  min_cut_reads=0)
        ex = SR.SplitExchange(len(lens))
        got = {}
        order = [SR.split_order(parts) for parts in plan]
        for f, _, _ in order:
            for p in f:
                r = run(p["seq"], first_read=p["first_read"], max_reads=p["max_reads"])
                assert r[2].res_num == p["max_reads"] and r[2].len_total_end < depth * lens[p["seq"]]
                ex.add(p["seq"], r[2].len_total_end)
                got[(p["seq"], p["first_read"])] = r
        ex.publish()
        for _, w, d in order:
            for p in w:
                got[(p["seq"], 0)] = run(p["seq"])
            for p in d:
                got[(p["seq"], p["first_read"])] = run(p["seq"], first_read=p["first_read"],
                                                       len_total_start=ex.prefix(p["seq"]))
        assert any(k[1] > 0 for k in got)   # something was split
        for k in range(len(lens)):
            parts = [got[key] for key in sorted(got) if key[0] == k]
            assert b"".join(p[0] for p in parts) == whole[k][0]
            assert b"".join(p[1] for p in parts) == whole[k][1]
            assert sum(p[2].res_num for p in parts) == whole[k][2].res_num
            assert sum(p[2].res_len_total for p in parts) == whole[k][2].res_len_total
            assert parts[-1][2].len_total_end == whole[k][2].len_total_end


def test_megabase_reads_equal_oracle(eng):
    """reads at the length limit (--length-mean 900000 --length-sd 0 --length-max 1000000): ~880 segments per read,
    backward coupling on every one of them; bytes and statistics must equal the oracle's"""
    from tests.golden_util import model_path
    from oracle import refrun as R
    okw = dict(len_mean=900000.0, len_sd=0.0, len_max=1000000, ratio=(39, 24, 36))
    contigs = R.synth_genome(31, [("big", 3000000)], n_runs=4, hp_plants=30)
    o = O.Oracle("qshmm", model_path("QSHMM-ONT.model"), **okw)
    o.rng_philox(17)
    o.set_sequence(contigs[0][1], 1)
    want_reads, want_maf, want_st = o.simulate_wgs(2.0)
    hm = capi.HostModel(capi.load(), capi.host_params("qshmm", **okw), model_path("QSHMM-ONT.model"))
    run = simulator.WgsRun(eng, hm, 2.0)
    reads, maf, st, text = run.simulate_sequence(contigs[0][1], 1, rng_mode=capi.RNG_PHILOX, seed=17)
    assert st.res_len_max > 850000
    assert reads == want_reads
    assert maf == want_maf
    assert text == O.format_stats(want_st, 1)


def test_long_read_roundtrip_properties_at_scale(eng):
    """60 Mbp device-generated contig, QSHMM-ONT 50 kb reads (the default bench workload's parameters), depth 3:
    every MAF block must re-derive its FASTQ read and its window of the genome; totals must match the statistics"""
    import ctypes as C
    from tests.golden_util import model_path
    okw = dict(len_mean=50000.0, len_sd=35000.0, len_max=1000000, ratio=(39, 24, 36))
    hm = capi.HostModel(capi.load(), capi.host_params("qshmm", **okw), model_path("QSHMM-ONT.model"))
    eng.set_model(hm)
    n = 60000000
    eng.set_synthetic_sequence(n, 1, 123)
    buf = (C.c_char * n)()
    eng.get_sequence_ascii(C.addressof(buf), n)
    genome = bytes(buf)
    reads, maf, st, _ = eng.simulate(3 * n, rng_mode=capi.RNG_PHILOX, seed=5)
    fq = reads.split(b"\n")
    blocks = maf.split(b"\n\n")[:-1]
    assert len(fq) // 4 == len(blocks) == st.res_num > 2000
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    total = ndel = nins = 0
    for k, blk in enumerate(blocks):
        l1, l2 = blk.split(b"\n")[1:3]
        f1, f2 = l1.split(), l2.split()
        off, wlen, refrow = int(f1[2]), int(f1[3]), f1[6]
        rlen, strand, readrow = int(f2[3]), f2[4], f2[6]
        assert len(refrow) == len(readrow)
        assert refrow.replace(b"-", b"") == genome[off:off + wlen]
        seq = fq[4 * k + 1]
        assert len(seq) == rlen == len(fq[4 * k + 3])
        rr = readrow.replace(b"-", b"")
        assert (rr if strand == b"+" else rr.translate(comp)[::-1]) == seq
        assert fq[4 * k] == b"@S1_%d" % (k + 1) and (strand == b"+") == (k % 2 == 0)
        total += rlen
        ndel += readrow.count(b"-")
        nins += refrow.count(b"-")
    assert total == st.res_len_total >= 3 * n
    assert ndel == st.res_del_num and nins == st.res_ins_num


@pytest.mark.parametrize("name", ["qs_rsii_basic", "qs_hp11_uniform", "qs_delheavy_uniform", "err_onthq_basic",
                                  "err_sequel_hiacc", "err_sequel_multipass"])
def test_one_segment_chain_chunks_equal_oracle(eng, name):
    """chain_chunk = 1 and seg_min_len = 1024: every segment's entry state comes from a backward coupling of its own
    (sticky chains: from the whole-read walk); bytes must still equal the oracle's"""
    c = Case(name)
    out, _ = c.run_oracle("philox")
    eng.set_option("seg_min_len", 1024)
    eng.set_option("chain_chunk", 1)
    try:
        res = run_case_on_gpu(c, eng, "philox")
    finally:
        eng.set_option("seg_min_len", 2048)
        eng.set_option("chain_chunk", 0)
    for i, ((reads, maf, st, text), o) in enumerate(zip(res, out), start=1):
        assert reads == o["reads"], "reads differ from the oracle, seq %d" % i
        assert maf == o["maf"], "maf differs from the oracle, seq %d" % i
        assert text == o["stats_text"]


def test_long_errhmm_reads_in_several_chunks_equal_oracle(eng):
    """errhmm reads of 150 k columns: five chain chunks each (coupling at 32 k-column boundaries), accuracies below
    and above the model's range included"""
    from tests.golden_util import model_path
    from oracle import refrun as R
    okw = dict(len_mean=150000.0, len_sd=0.0, len_max=1000000)
    contigs = R.synth_genome(41, [("e1", 2000000)], n_runs=3, hp_plants=20)
    o = O.Oracle("errhmm", model_path("ERRHMM-ONT-HQ.model"), **okw)
    o.rng_philox(23)
    o.set_sequence(contigs[0][1], 1)
    want_reads, want_maf, want_st = o.simulate_wgs(3.0)
    hm = capi.HostModel(capi.load(), capi.host_params("errhmm", **okw), model_path("ERRHMM-ONT-HQ.model"))
    run = simulator.WgsRun(eng, hm, 3.0)
    reads, maf, st, text = run.simulate_sequence(contigs[0][1], 1, rng_mode=capi.RNG_PHILOX, seed=23)
    assert st.res_num >= 40
    assert reads == want_reads
    assert maf == want_maf
    assert text == O.format_stats(want_st, 1)
