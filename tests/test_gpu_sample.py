"""GPU: --method sample (simulate_by_sample, pbsim.cpp:1694) through the C ABI — the reference's bytes in replay
mode, the oracle's in PHILOX mode, for every batch size (the copies of a pool entry are a chain that one GPU thread
walks; batches cut the pool passes anywhere between entries)."""
import numpy as np
import pytest

from oracle import oracle as O
from pbsim_b200 import capi, simulator
from tests.golden_util import SampleCase, sample_case_names

pytestmark = pytest.mark.gpu


def _oracle(c, rng):
    o = O.Oracle("sample", None, **c.okw)
    (o.rng_glibc if rng == "glibc" else o.rng_philox)(c.seed)
    if c.okw.get("hp_del_bias", 1.0) != 1.0:
        o.hp_bias_prepass([s for _, s in c.contigs])
    out = []
    for i, (_, s) in enumerate(c.contigs, start=1):
        o.set_sequence(s, i)
        reads, maf, st = o.simulate_sample(c.depth, c.pool)
        out.append(dict(reads=reads, maf=maf, text=O.format_stats(st, i), n=len(o.readinfo())))
    return out


def _run(c, spec=1):
    hm = capi.HostModel(capi.load(), capi.host_params("sample", **c.okw), None)
    eng = simulator.Engine(0)
    eng.set_option("pipeline", 0)
    eng.set_option("sample_spec", spec)  # 1: all copies at once assuming full-length reads, wrong tails redone as chains
    run = simulator.WgsRun(eng, hm, c.depth, c.okw.get("hp_del_bias", 1.0))
    eng.set_pool(c.pool)
    if c.okw.get("hp_del_bias", 1.0) != 1.0:
        run.prepass(c.contigs)
    return eng, run


@pytest.mark.parametrize("spec", [0, 1])
@pytest.mark.parametrize("batch_reads", [0, 1, 7])
@pytest.mark.parametrize("name", sample_case_names())
def test_replay_reproduces_reference(name, batch_reads, spec):
    c = SampleCase(name)
    ref = _oracle(c, "glibc")
    draws = O.glibc_rand(c.seed, c.ndraws)
    starts = np.concatenate([[0], c.marks[:-1]]).astype(np.int64)
    eng, run = _run(c, spec)
    pos = 0
    for i, (_, s) in enumerate(c.contigs, start=1):
        n = ref[i - 1]["n"]
        end = int(starts[pos + n]) if pos + n < len(starts) else c.ndraws
        reads, maf, st, text = run.simulate_sequence(s, i, rng_mode=capi.RNG_REPLAY, replay_draws=draws[:end],
                                                     replay_starts=starts[pos:pos + n], batch_reads=batch_reads)
        pos += n
        assert reads == c.reads(i), "FASTQ differs from the reference (seq %d)" % i
        assert maf == c.maf(i), "MAF differs from the reference (seq %d)" % i
        assert text == c.stats_blocks[i]


@pytest.mark.parametrize("spec", [0, 1])
@pytest.mark.parametrize("batch_reads", [0, 5])
@pytest.mark.parametrize("name", sample_case_names())
def test_philox_equals_oracle(name, batch_reads, spec):
    c = SampleCase(name)
    ref = _oracle(c, "philox")
    eng, run = _run(c, spec)
    for i, (_, s) in enumerate(c.contigs, start=1):
        reads, maf, st, text = run.simulate_sequence(s, i, rng_mode=capi.RNG_PHILOX, seed=c.seed, batch_reads=batch_reads)
        assert reads == ref[i - 1]["reads"]
        assert maf == ref[i - 1]["maf"]
        assert text == ref[i - 1]["text"]


def test_pipelined_and_device_delivery_match():
    c = SampleCase("sample_quirks")
    ref = _oracle(c, "philox")
    from tests.test_gpu_parity import _device_bytes
    for mode in ("pipeline", "device"):
        eng, run = _run(c)
        eng.set_option("pipeline", 2 if mode == "pipeline" else 0)
        for i, (_, s) in enumerate(c.contigs, start=1):
            if mode == "pipeline":
                reads, maf, st, text = run.simulate_sequence(s, i, rng_mode=capi.RNG_PHILOX, seed=c.seed, batch_reads=9)
            else:
                bias = list(run.base_bias)
                eng.set_sequence(s, i, bias)
                run.hp11_running += eng.hpfreq()[11]
                bias[0] = float(np.array([run.hp11_running], dtype=np.int64).view(np.float64)[0])
                eng.update_bias(bias)
                eng.begin(int(c.depth * len(s)), rng_mode=capi.RNG_PHILOX, seed=c.seed, batch_reads=11)
                rs, ms = [], []
                while True:
                    ch = eng.next_chunk(device=True)
                    if ch is None:
                        break
                    rs.append(_device_bytes(ch.reads, ch.reads_bytes))
                    ms.append(_device_bytes(ch.maf, ch.maf_bytes))
                eng.end()
                reads, maf = b"".join(rs), b"".join(ms)
            assert reads == ref[i - 1]["reads"], mode
            assert maf == ref[i - 1]["maf"], mode


def test_misuse_fails_loudly():
    c = SampleCase("sample_basic")
    hm = capi.HostModel(capi.load(), capi.host_params("sample", **c.okw), None)
    eng = simulator.Engine(0)
    eng.set_model(hm)
    eng.set_sequence(c.contigs[0][1], 1, [0.0] + [1.0] * 10 + [0.0])
    with pytest.raises(simulator.EngineError, match="set_pool"):
        eng.begin(1000)
    with pytest.raises(simulator.EngineError, match="at least 2 reads"):
        eng.set_pool([b"IIII"])
    with pytest.raises(simulator.EngineError, match="quality character"):
        eng.set_pool([b"II II", b"IIII"])
    eng.set_pool(c.pool)
    with pytest.raises(simulator.EngineError, match="read ranges"):
        eng.begin(1000, first_read=5)


@pytest.mark.parametrize("spec", [0, 1])
@pytest.mark.parametrize("glen,depth,lens", [
    (2000, 1.5, [300, 250, 200, 150, 100]),      # quota is a multiple of the pool: sample_interval = 1 (:1721)
    (5000, 3.3, [400, 350, 120]),                # n * 0.5 truncates to 1: the interval clamp (:1727)
    (700, 9.0, [900, 650, 300, 800, 120, 101]),  # entries longer than the sequence: offset 0, len = glen (:1758)
    (4000, 0.2, [500, 400, 300, 200]),           # sample_num = 0: the first pool pass copies only the extras
    (3000, 40.0, [150, 140, 130, 120, 110, 100, 160, 170]),  # many copies per entry, groups straddle batches
])
def test_schedule_corners_equal_oracle(glen, depth, lens, spec):
    rng = np.random.default_rng(glen)
    genome = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), glen))
    pool = [bytes(rng.integers(33 + 5, 33 + 20, n).astype(np.uint8)) for n in lens]
    okw = dict(ratio=(20, 30, 50), len_min=100, len_max=2500)
    o = O.Oracle("sample", None, **okw)
    o.rng_philox(9)
    o.set_sequence(genome, 1)
    oreads, omaf, ost = o.simulate_sample(depth, pool)
    hm = capi.HostModel(capi.load(), capi.host_params("sample", **okw), None)
    eng = simulator.Engine(0)
    eng.set_option("pipeline", 0)
    eng.set_option("sample_spec", spec)
    run = simulator.WgsRun(eng, hm, depth)
    eng.set_pool(pool)
    reads, maf, st, text = run.simulate_sequence(genome, 1, rng_mode=capi.RNG_PHILOX, seed=9, batch_reads=3)
    assert reads == oreads and maf == omaf
    assert text == O.format_stats(ost, 1)


@pytest.mark.parametrize("segments", [1, 0])
@pytest.mark.parametrize("glen,depth,lens,ratio", [
    # long pool entries: the speculative pass runs them through the error pass (k_sim_seg<true>) in segments of 1024
    # positions and k_find_end; lengths around the segment size, a multiple of it, and one entry longer than the sequence
    (60000, 6.0, [5000, 3072, 2048, 2049, 4095, 9000, 1500, 300, 70000], (20, 50, 30)),  # insertion-rich: reads end first
    (60000, 6.0, [5000, 3072, 2048, 2049, 4095, 9000, 1500, 300], (20, 20, 60)),   # deletion-rich: windows end first, chains redone
    (30000, 12.0, [8000, 6000, 2500], (6, 55, 39)),                                # several copies per entry
])
def test_long_entries_on_segments_equal_oracle(glen, depth, lens, ratio, segments):
    rng = np.random.default_rng(glen + len(lens))
    g = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), glen).tobytes())
    # exceptional blocks: homopolymers of 11 and more (no deletion may follow them), N runs, IUPAC codes, lower case
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
    hm = capi.HostModel(capi.load(), capi.host_params("sample", **okw), None)
    eng = simulator.Engine(0)
    eng.set_option("pipeline", 0)
    eng.set_option("segments", segments)
    run = simulator.WgsRun(eng, hm, depth)
    eng.set_pool(pool)
    reads, maf, st, text = run.simulate_sequence(genome, 1, rng_mode=capi.RNG_PHILOX, seed=21)
    assert reads == oreads and maf == omaf
    assert text == O.format_stats(ost, 1)
    assert st.res_len_max >= 8000
