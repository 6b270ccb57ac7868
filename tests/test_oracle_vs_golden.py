"""Pins the CPU oracle (oracle/pbsim_oracle.c) against outputs of the UNMODIFIED reference binary
captured in tests/golden/ (oracle/make_golden.py).  Byte-exact: FASTQ / SAM, MAF, the stderr
stats block, the number of rand() calls and the per-(read, pass) draw offsets."""
import numpy as np
import pytest

from oracle import oracle as O

from tests.golden_util import Case, SampleCase, SetCase, case_names, sample_case_names, set_case_names


@pytest.mark.parametrize("name", case_names())
def test_oracle_reproduces_reference(name):
    c = Case(name)
    out, o = c.run_oracle("glibc")
    starts = []
    for i, res in enumerate(out, start=1):
        assert res["reads"] == c.reads(i), "reads differ, seq %d" % i
        assert res["maf"] == c.maf(i), "maf differs, seq %d" % i
        assert res["stats_text"] == c.stats_blocks[i], "stats differ, seq %d" % i
        starts.append(res["info"]["draw_start"])
    starts = np.concatenate(starts)
    assert o.draws_consumed() == c.ndraws
    assert len(starts) == len(c.marks)
    assert np.array_equal(starts[1:], c.marks[:-1])
    assert c.marks[-1] == c.ndraws


@pytest.mark.parametrize("name", ["qs_rsii_quirks", "err_sequel_multipass"])
def test_oracle_replay_equals_glibc(name):
    c = Case(name)
    out, o = c.run_oracle("glibc")
    log = o.draw_log()
    out2, _ = c.run_oracle("replay", log=log)
    for a, b in zip(out, out2):
        assert a["reads"] == b["reads"] and a["maf"] == b["maf"]


def test_philox_mode_is_deterministic_and_differs_from_glibc():
    c = Case("qs_rsii_basic")
    a, _ = c.run_oracle("philox")
    b, _ = c.run_oracle("philox")
    g, _ = c.run_oracle("glibc")
    assert a[0]["reads"] == b[0]["reads"] and a[0]["maf"] == b[0]["maf"]
    assert a[0]["reads"] != g[0]["reads"]


@pytest.mark.parametrize("name", set_case_names())
def test_oracle_reproduces_reference_transcript_and_template_runs(name):
    """--strategy trans / templ (pbsim.cpp:2419, :3055, :4114, :4807): same bytes, stats block and rand() count"""
    c = SetCase(name)
    res, o = c.run_oracle("glibc")
    assert res["reads"] == c.reads()
    assert res["maf"] == c.maf()
    assert res["stats_text"] == c.stats_text
    assert o.draws_consumed() == c.ndraws
    starts = res["info"]["draw_start"]
    assert len(starts) == len(c.marks)
    assert np.array_equal(starts[1:], c.marks[:-1])


def test_start_position_table_known_answers():
    """prob2ssp (pbsim.cpp:2504-2528): rank 1 puts 62.6 % of the reads at the 5' end, higher ranks fewer"""
    ends, mod = O.ssp_table(14)
    assert mod[1:].tolist() == [1000] * 14
    assert all(np.all(np.diff(ends[r][ends[r] >= 0]) >= 0) for r in range(1, 15))
    assert ends[1][0] == 626 and ends[2][0] == 458 and ends[10][0] == 310  # 1 / sum_{j<=21} j^-2 = 0.6256


@pytest.mark.parametrize("name", sample_case_names())
def test_oracle_reproduces_reference_sample_method(name):
    """--method sample (pbsim.cpp:1694; get_sample_inf's filter :1214-1275 restated in oracle.sample_pool): the
    specification of the engine's sample method, pinned byte for byte — including the quality buffer that is cut at
    the end of every read, which makes copy i+1 of a pool entry as long as copy i's read"""
    c = SampleCase(name)
    o = O.Oracle("sample", None, **c.okw)
    o.rng_glibc(c.seed)
    if c.okw.get("hp_del_bias", 1.0) != 1.0:
        o.hp_bias_prepass([s for _, s in c.contigs])
    starts = []
    for i, (_, s) in enumerate(c.contigs, start=1):
        o.set_sequence(s, i)
        reads, maf, st = o.simulate_sample(c.depth, c.pool)
        assert reads == c.reads(i), "reads differ, seq %d" % i
        assert maf == c.maf(i), "maf differs, seq %d" % i
        assert O.format_stats(st, i) == c.stats_blocks[i]
        starts.append(o.readinfo()["draw_start"])
    assert o.draws_consumed() == c.ndraws
    starts = np.concatenate(starts)
    assert np.array_equal(starts[1:], c.marks[:-1])
