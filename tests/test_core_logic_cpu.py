"""CPU checks of the product's host front end (pbsim_b200/csrc/host_model.cpp, model_image.hpp) and of
the engine's pass-1 core (sim_core.cuh, compiled as plain C++ by tests/hostsim) against the oracle and the
reference's golden outputs.  The GPU tests (-m gpu) then check the CUDA engine end to end."""
import numpy as np
import pytest

from oracle import oracle as O
from pbsim_b200 import capi
from tests import hostsim_util as H
from tests.golden_util import Case, case_names


def load_product_model(c):
    return capi.HostModel(H.lib(), capi.host_params(c.method, **c.okw), c.model)


@pytest.mark.parametrize("name", case_names())
def test_host_tables_equal_oracle_tables(name):
    """KAT: quantised tables of the product's builder == the oracle's (which reproduces the reference)."""
    c = Case(name)
    _compare_tables(c.new_oracle(), load_product_model(c), c.method)


@pytest.mark.parametrize("k", range(40))
def test_host_tables_equal_oracle_tables_on_random_options(k):
    """the same for 40 random option sets (accuracy 0.70 .. 1.00, fixed lengths, all seven models): the configurations on
    which tests/test_oracle_vs_reference_live_cpu.py pins the oracle to the reference"""
    from tests.golden_util import model_path
    from tests.test_oracle_vs_reference_live_cpu import _config_wide
    cfg = _config_wide(k)
    mp = model_path(cfg["model"])
    o = O.Oracle(cfg["method"], mp, **cfg["okw"])
    hm = capi.HostModel(H.lib(), capi.host_params(cfg["method"], **cfg["okw"]), mp)
    _compare_tables(o, hm, cfg["method"])


def _compare_tables(o, hm, method):
    v = hm.view
    t, n = o.table(0)
    assert v.len_rand_value == n and np.array_equal(np.ctypeslib.as_array(v.prob2len, shape=(n,)), t)
    t, n = o.table(1)
    assert v.accuracy_rand_value == n and np.array_equal(np.ctypeslib.as_array(v.prob2accuracy, shape=(n,)), t)
    s, i, d = o.thresholds()
    assert list(v.sub_thre) == s.tolist() and list(v.ins_thre) == i.tolist() and list(v.del_thre) == d.tolist()
    amin, amax, lo, hi = o.model_range()
    assert (v.acc_lo, v.acc_hi) == (lo, hi)
    err = method == "errhmm"
    if err:
        assert (v.model_acc_min, v.model_acc_max) == (amin, amax)
    checked = 0
    for a in range(lo, hi + 1):
        r = v.rows[a]
        assert bool(r.exists) == o.model_exists(a)
        if not r.exists:
            if not err:
                t, n = o.table(5, a, cap=1001)
                assert r.freq_mod == n and np.array_equal(np.ctypeslib.as_array(r.freq, shape=(n,)), t)
                checked += 1
            continue
        res = r.resolution
        t, n = o.table(2, a, cap=1001)
        assert r.init_mod == n and np.array_equal(np.ctypeslib.as_array(r.init, shape=(n,)), t)
        for st in range(1, r.nstates + 1):
            t, n = o.table(3, a, st, cap=1001)
            assert r.emis_mod[st] == n, (a, st)
            if 0 < n <= res:
                assert np.array_equal(np.ctypeslib.as_array(r.emis, shape=((r.nstates + 1) * res,))[st * res:st * res + n], t)
            t, n = o.table(4, a, st, cap=1001)
            assert r.tran_mod[st] == n, (a, st)
            if 0 < n <= res:
                assert np.array_equal(np.ctypeslib.as_array(r.tran, shape=((r.nstates + 1) * res,))[st * res:st * res + n], t)
            if err:
                assert r.emis_del[st] == o.emis2del(a, st)
            checked += 1
    assert checked > 0


def test_hp_del_bias_front_end_matches_oracle():
    c = Case("qs_ont_hpbias")
    out, o = c.run_oracle("glibc")
    L = H.lib()
    # hpfreq over all sequences, as main's prepass accumulates it (pbsim.cpp:678-685)
    hpfreq = np.zeros(12, dtype=np.int64)
    o2 = c.new_oracle()
    for i, (_, s) in enumerate(c.contigs, start=1):
        o2.set_sequence(s, i)
        hp = o2.hp()
        up = np.frombuffer(o2.seq_upper(), dtype=np.uint8)
        for h in range(1, 12):
            hpfreq[h] += int(np.count_nonzero(hp == h))
    bias = capi.hp_del_bias(L, c.okw["hp_del_bias"], hpfreq)
    # oracle bias during the first sequence: [1..10] identical; [0] aliases the running hp==11 count
    assert np.allclose(bias[1:11], out[0]["bias"][1:11], rtol=0, atol=0)


def _replay_case(c, rng):
    """Run hostsim over all sequences of a case; returns per-sequence (reads, maf, subreads)."""
    out, o = c.run_oracle("glibc" if rng == "replay" else "philox")
    log = o.draw_log() if rng == "replay" else None
    hm = load_product_model(c)
    res = []
    o2 = c.new_oracle()
    if c.okw.get("hp_del_bias", 1.0) != 1.0:
        o2.hp_bias_prepass([s for _, s in c.contigs])
    cursor = 0
    for i, (_, s) in enumerate(c.contigs, start=1):
        o2.set_sequence(s, i)
        up, hp, bias = o2.seq_upper(), o2.hp(), o2.bias()
        quota = int(c.depth * len(s))
        if rng == "replay":
            sub = H.run(hm, up, hp, i, bias, capi.RNG_REPLAY, 0, log[cursor:], quota)
            for sr in sub:
                sr["draw_start"] += cursor
            cursor = out[i - 1]["draws_end"]
        else:
            sub = H.run(hm, up, hp, i, bias, capi.RNG_PHILOX, c.seed, None, quota)
        reads, maf = H.records_from_events(hm, sub, up, i)
        res.append((reads, maf, sub))
    return out, res


@pytest.mark.parametrize("name", case_names())
def test_core_replay_reproduces_reference(name):
    """pass-1 core fed the reference's own draws + the pass-2 specification == reference bytes."""
    c = Case(name)
    out, res = _replay_case(c, "replay")
    for i, ((reads, maf, sub), oref) in enumerate(zip(res, out), start=1):
        assert reads == c.reads(i), "reads differ, seq %d" % i
        assert maf == c.maf(i), "maf differs, seq %d" % i
        info = oref["info"]
        assert [s["draw_start"] for s in sub] == info["draw_start"].tolist()
        assert [s["nsub"] for s in sub] == info["nsub"].tolist()
        assert [s["ndel"] for s in sub] == info["ndel"].tolist()
        assert np.array_equal(np.array([s["accuracy"] for s in sub]), info["accuracy"])
        assert not any(s["overflow"] for s in sub)


@pytest.mark.parametrize("name", ["qs_rsii_quirks", "qs_ont_hpbias", "err_onthq_basic", "err_sequel_hiacc",
                                  "qs_rsii_multipass", "qs_delheavy_uniform", "qs_delheavy_bias", "qs_rsii_basic"])
def test_core_philox_equals_oracle_philox(name):
    c = Case(name)
    out, res = _replay_case(c, "philox")
    for (reads, maf, sub), oref in zip(res, out):
        assert reads == oref["reads"]
        assert maf == oref["maf"]


def test_fast_path_equals_generic_path():
    """qshmm_simulate_fast (what the GPU runs for PHILOX reads that never touch the genome in pass 1) expands to
    the same records as the generic per-draw path"""
    for name in ("qs_rsii_basic", "qs_delheavy_uniform"):
        c = Case(name)
        H.lib().hostsim_use_fast(1)
        _, fast = _replay_case(c, "philox")
        H.lib().hostsim_use_fast(0)
        try:
            _, gen = _replay_case(c, "philox")
        finally:
            H.lib().hostsim_use_fast(1)
        for (r1, m1, s1), (r2, m2, s2) in zip(fast, gen):
            assert r1 == r2 and m1 == m2
            assert [x["nins"] for x in s1] == [x["nins"] for x in s2]
            assert [x["ndel"] for x in s1] == [x["ndel"] for x in s2]


@pytest.mark.parametrize("name", ["qs_rsii_basic", "qs_rsii_quirks", "qs_delheavy_uniform", "qs_rsii_multipass",
                                  "qs_rsii_fixedlen", "qs_hp11_uniform", "err_onthq_basic", "err_sequel_multipass",
                                  "err_sequel_hiacc"])
def test_segment_parallel_pass1_equals_oracle(name):
    """segment-parallel pass 1 (backward coupling + unbounded segments + find_end) gives the oracle's bytes"""
    c = Case(name)
    L = H.lib()
    L.hostsim_seg_reads.restype = L.hostsim_seg_fallbacks.restype = __import__("ctypes").c_long
    L.hostsim_use_segments(1, 1025)
    try:
        out, res = _replay_case(c, "philox")
    finally:
        L.hostsim_use_segments(0, 2048)
    for (reads, maf, sub), oref in zip(res, out):
        assert reads == oref["reads"]
        assert maf == oref["maf"]
        assert np.array_equal(np.array([s["accuracy"] for s in sub]), oref["info"]["accuracy"])
        assert [s["nins"] for s in sub] == oref["info"]["nins"].tolist()
    if name != "qs_rsii_fixedlen":
        assert L.hostsim_seg_reads() > 0


def test_start_position_table_equals_oracle():
    """pbsim_host_ssp_table (the product's builder of prob2ssp, pbsim.cpp:2504-2528) against the oracle's, for every
    rank a 999 kb transcript can reach"""
    import ctypes as C
    from pbsim_b200 import capi
    L = capi.load()
    rank_max = 999
    ends = np.zeros((rank_max + 1) * 21, dtype=np.uint16)
    mod = np.zeros(rank_max + 1, dtype=np.uint16)
    L.pbsim_host_ssp_table(rank_max, ends.ctypes.data, mod.ctypes.data)
    o_ends, o_mod = O.ssp_table(rank_max)
    ends = ends.reshape(rank_max + 1, 21).astype(np.int64)
    ends[ends == 0xFFFF] = -1
    assert np.array_equal(ends[1:], o_ends[1:])
    assert np.array_equal(mod[1:], o_mod[1:])


@pytest.mark.parametrize("chunk", [1, 2, 32])
@pytest.mark.parametrize("name", ["qs_rsii_basic", "qs_hp11_uniform", "err_onthq_basic", "err_sequel_hiacc"])
def test_chain_chunks_give_the_oracle_bytes(name, chunk):
    """the states in front of the segments as k_chain_chunk recovers them (coupling at a chunk's first position,
    chain-only walk through the chunk), run on the host with chunks of 1, 2 and 32 segments"""
    c = Case(name)
    L = H.lib()
    L.hostsim_use_segments(1, 1025)
    L.hostsim_set_chain_chunk(chunk)
    try:
        out, res = _replay_case(c, "philox")
    finally:
        L.hostsim_use_segments(0, 2048)
        L.hostsim_set_chain_chunk(0)
    for (reads, maf, sub), oref in zip(res, out):
        assert reads == oref["reads"]
        assert maf == oref["maf"]


# ---- --method sample: the engine's schedule (sample_plan.hpp) and per-read core (sample_simulate) on the host

def _sample_case(c, rng, batch_reads=0):
    from tests.golden_util import SampleCase  # noqa: F401
    o = O.Oracle("sample", None, **c.okw)
    if rng == "replay":
        o.rng_glibc(c.seed)
    else:
        o.rng_philox(c.seed)
    if c.okw.get("hp_del_bias", 1.0) != 1.0:
        o.hp_bias_prepass([s for _, s in c.contigs])
    hm = capi.HostModel(H.lib(), capi.host_params("sample", **c.okw), None)
    out = []
    cursor = 0
    for i, (_, s) in enumerate(c.contigs, start=1):
        o.set_sequence(s, i)
        up, hp, bias = o.seq_upper(), o.hp(), o.bias()
        oreads, omaf, ost = o.simulate_sample(c.depth, c.pool)
        info = o.readinfo()
        quota = int(c.depth * len(s))
        if rng == "replay":
            log = o.draw_log()
            sub = H.run(hm, up, hp, i, bias, capi.RNG_REPLAY, 0, log[cursor:], quota, pool=c.pool, batch_reads=batch_reads)
            for sr in sub:
                sr["draw_start"] += cursor
            cursor = o.draws_consumed()
        else:
            sub = H.run(hm, up, hp, i, bias, capi.RNG_PHILOX, c.seed, None, quota, pool=c.pool, batch_reads=batch_reads)
        reads, maf = H.records_from_events(hm, sub, up, i)
        out.append((reads, maf, sub, oreads, omaf, info))
    return out


@pytest.mark.parametrize("batch_reads", [0, 7])
@pytest.mark.parametrize("name", ["sample_basic", "sample_quirks"])
def test_sample_core_replay_reproduces_reference(name, batch_reads):
    from tests.golden_util import SampleCase
    c = SampleCase(name)
    for i, (reads, maf, sub, oreads, omaf, info) in enumerate(_sample_case(c, "replay", batch_reads), start=1):
        assert reads == c.reads(i), "reads differ, seq %d" % i
        assert maf == c.maf(i), "maf differs, seq %d" % i
        assert [s["draw_start"] for s in sub] == info["draw_start"].tolist()
        assert np.array_equal(np.array([s["accuracy"] for s in sub]), info["accuracy"])
        assert not any(s["overflow"] for s in sub)


@pytest.mark.parametrize("name", ["sample_basic", "sample_quirks"])
def test_sample_core_philox_equals_oracle_philox(name):
    from tests.golden_util import SampleCase
    c = SampleCase(name)
    for reads, maf, sub, oreads, omaf, info in _sample_case(c, "philox", 5):
        assert reads == oreads
        assert maf == omaf
        assert [s["nins"] for s in sub] == info["nins"].tolist()


@pytest.mark.parametrize("name", ["sample_basic", "sample_quirks"])
def test_sample_filter_equals_reference(name):
    """pbsim_host_sample_filter (get_sample_inf, pbsim.cpp:1155-1330): the pool equals the oracle's restatement and the
    statistics are the block the reference printed"""
    from tests.golden_util import SampleCase
    c = SampleCase(name)
    pool, st = capi.sample_filter(H.lib(), c.sample_fastq, **c.meta["pool_kwargs"])
    assert pool == c.pool
    assert capi.format_sample_stats(st, "sample.fq") in c.stderr
    # a last record without line feed is not counted; nothing in range is the reference's error
    cut, _ = capi.sample_filter(H.lib(), c.sample_fastq[:-1], **c.meta["pool_kwargs"])
    assert cut == c.pool[:-1] or len(c.pool[-1]) and cut == c.pool[:len(cut)] and len(cut) >= len(c.pool) - 1
    with pytest.raises(RuntimeError, match="no sample in the valid range"):
        capi.sample_filter(H.lib(), c.sample_fastq, len_min=999999, len_max=1000000)


@pytest.mark.parametrize("glen,depth,lens", [
    (2000, 1.5, [300, 250, 200, 150, 100]),      # quota is a multiple of the pool: sample_interval = 1 (:1721)
    (5000, 3.3, [400, 350, 120]),                # n * 0.5 truncates to 1: the interval clamp (:1727)
    (700, 9.0, [900, 650, 300, 800, 120, 101]),  # entries longer than the sequence: offset 0, len = glen (:1758)
    (4000, 0.2, [500, 400, 300, 200]),           # sample_num = 0: the first pool pass already copies only the extras
    (3000, 40.0, [150, 140, 130, 120, 110, 100, 160, 170]),  # many copies per entry, several batches per group
])
def test_sample_schedule_edge_cases_equal_oracle(glen, depth, lens):
    """the group schedule (sample_plan.hpp) against the oracle's sequential loop where the reference's arithmetic has
    corners; PHILOX draws, batches of 3 reads so that groups straddle batches"""
    rng = np.random.default_rng(glen)
    genome = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), glen))
    pool = [bytes(rng.integers(33 + 5, 33 + 20, n).astype(np.uint8)) for n in lens]
    okw = dict(ratio=(20, 30, 50), len_min=100, len_max=2500)
    o = O.Oracle("sample", None, **okw)
    o.rng_philox(9)
    o.set_sequence(genome, 1)
    oreads, omaf, ost = o.simulate_sample(depth, pool)
    hm = capi.HostModel(H.lib(), capi.host_params("sample", **okw), None)
    sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_PHILOX, 9, None, int(depth * glen), pool=pool,
                batch_reads=3)
    reads, maf = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert reads == oreads and maf == omaf
    assert len(sub) == ost.res_num


@pytest.mark.parametrize("glen,depth,lens,ratio", [
    (60000, 6.0, [5000, 3072, 2048, 2049, 4095, 9000, 1500, 300, 70000], (20, 50, 30)),  # insertion-rich: reads end first
    (60000, 6.0, [5000, 3072, 2048, 2049, 4095, 9000, 1500, 300], (20, 20, 60)),         # deletion-rich: windows end first
    (30000, 12.0, [8000, 6000, 2500], (6, 55, 39)),
])
def test_sample_long_entries_on_segments_equal_oracle(glen, depth, lens, ratio):
    """--method sample, long pool entries on the segment path as the kernels run it (hostsim run_segmented_sample: the
    qualities of a segment's positions from the pool entry, qshmm's error pass, the read ended where its window is
    used up or where it is as long as its quality string: qshmm_finish_segmented / qshmm_walk_tile with p_left), on a
    genome with homopolymers of 11 and more, N runs, IUPAC codes and lower case; the oracle's sequential loop decides"""
    rng = np.random.default_rng(glen + len(lens))
    g = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), glen).tobytes())
    for k, at in enumerate(rng.integers(100, glen - 100, 14)):
        run = [b"A" * 12, b"T" * 15, b"NNNNN", b"R", b"c" * 11, b"G" * 30, b"y"][k % 7]
        g[at:at + len(run)] = run
    genome = bytes(g)
    pool = [bytes(rng.integers(33 + 3, 33 + 25, n).astype(np.uint8)) for n in lens]
    okw = dict(ratio=ratio, len_min=100, len_max=100000)
    o = O.Oracle("sample", None, **okw)
    o.rng_philox(21)
    o.set_sequence(genome, 1)
    oreads, omaf, ost = o.simulate_sample(depth, pool)
    hm = capi.HostModel(H.lib(), capi.host_params("sample", **okw), None)
    L = H.lib()
    L.hostsim_use_segments(1, 2048)
    try:
        before = L.hostsim_seg_reads()
        sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_PHILOX, 21, None, int(depth * glen), pool=pool)
        assert L.hostsim_seg_reads() - before >= 10 and L.hostsim_seg_fallbacks() == 0
    finally:
        L.hostsim_use_segments(0, 2048)
    reads, maf = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert reads == oreads and maf == omaf
    assert len(sub) == ost.res_num


@pytest.mark.parametrize("strategy", ["trans", "templ"])
def test_set_planners_equal_oracle_on_long_sequences(strategy):
    """plan_read_trans / plan_read_templ (what k_plan runs per read of a sequence set) against the oracle's read plans
    in PHILOX mode, on transcripts of up to 900 kb (start-position tables of ranks up to 900, pbsim.cpp:2504-2528,
    :2842-2866) and with windows clipped at the transcript's end"""
    import ctypes as C
    from tests.golden_util import model_path
    rng = np.random.default_rng(77)
    lens = [int(x) for x in rng.integers(300, 900000, 14)] + [999000, 1000, 1001, 21, 22]
    seqset = [("T%d" % i, int(rng.integers(0, 6)), int(rng.integers(0, 6)),
               np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes()) for i, n in enumerate(lens)]
    okw = dict(len_mean=30000.0, len_sd=25000.0, len_max=1000000, accuracy_mean=0.9, accuracy_mean_set=True)
    o = O.Oracle("qshmm", model_path("QSHMM-RSII.model"), **okw)
    o.rng_philox(5)
    o.simulate_set(strategy, seqset)
    info = o.readinfo()
    hm = capi.HostModel(H.lib(), capi.host_params("qshmm", **okw), model_path("QSHMM-RSII.model"))
    L = H.lib()
    L.hostsim_plan_set.restype = C.c_long
    L.hostsim_plan_set.argtypes = [C.POINTER(capi.Model), C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                   C.c_int, C.c_void_p, C.c_long]
    tlen = np.array(lens, dtype=np.int64)
    plus = np.array([x[1] for x in seqset], dtype=np.int32)
    minus = np.array([x[2] for x in seqset], dtype=np.int32)
    out = np.zeros((len(info) + 8, 4), dtype=np.int64)
    n = L.hostsim_plan_set(hm.ptr, capi.STRATEGY_TRANS if strategy == "trans" else capi.STRATEGY_TEMPL, len(lens),
                           tlen.ctypes.data, plus.ctypes.data, minus.ctypes.data, 5, 1000, out.ctypes.data, len(out))
    assert n == len(info) and n > 10
    out = out[:n]
    assert np.array_equal(out[:, 0], info["offset"])
    assert np.array_equal(out[:, 1], info["wlen"])
    assert np.array_equal(out[:, 2], info["acc"])
    assert np.array_equal(out[:, 3], (info["strand"] == ord("-")).astype(np.int64))


@pytest.mark.parametrize("k", range(12))
def test_hp_del_bias_front_end_equals_oracle_on_random_genomes(k):
    """pbsim_host_hp_del_bias (main :673-697, with its long += double truncation) on random multi-contig genomes and
    bias options: cells 1..10 equal the oracle's after its prepass; cell 0 is the aliased hpfreq[11] counter"""
    from oracle import refrun as R
    rng = np.random.default_rng(31000 + k)
    contigs = R.synth_genome(600 + k, [("c%d" % t, int(rng.integers(200, 30000))) for t in range(int(rng.integers(1, 6)))],
                             n_runs=int(rng.integers(0, 4)), hp_plants=int(rng.integers(0, 60)), iupac=int(rng.integers(0, 4)),
                             lowercase_frac=float(rng.choice([0.0, 0.3])), long_runs=(11, 12, 13, 26) if k % 2 else ())
    opt = float(rng.choice([1.5, 2.0, 3.0, 7.25, 10.0]))
    o = O.Oracle("qshmm", __import__("tests.golden_util", fromlist=["model_path"]).model_path("QSHMM-RSII.model"),
                 hp_del_bias=opt)
    o.hp_bias_prepass([s for _, s in contigs])
    hpfreq = np.zeros(12, dtype=np.int64)
    for i, (_, s) in enumerate(contigs, start=1):
        o.set_sequence(s, i)
        hp = o.hp()
        for h in range(1, 12):
            hpfreq[h] += int(np.count_nonzero(hp == h))
    bias = capi.hp_del_bias(H.lib(), opt, hpfreq)
    assert bias[1:11] == o.bias()[1:11].tolist()
    assert bias[11] == 0.0
    assert np.array([bias[0]]).view(np.int64)[0] == hpfreq[11]
