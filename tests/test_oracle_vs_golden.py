"""Pins the CPU oracle (oracle/pbsim_oracle.c) against outputs of the UNMODIFIED reference binary
captured in tests/golden/ (oracle/make_golden.py).  Byte-exact: FASTQ / SAM, MAF, the stderr
stats block, the number of rand() calls and the per-(read, pass) draw offsets."""
import numpy as np
import pytest

from tests.golden_util import Case, case_names


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
