"""PHILOX mode cannot reproduce the reference's libc rand() stream; north star (b) asks for read-length,
accuracy, error-type and quality distributions that match the reference statistically.  The fixtures under
tests/golden/stats/ are reductions of UNMODIFIED reference runs (oracle/make_golden.py); tolerances are
stated in tests/stats_util.compare.  CPU test: the oracle's Philox mode (which the GPU engine matches bit for
bit, tests/test_gpu_parity.py); GPU test: the engine itself."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refrun as R
from tests import stats_util as SU
from tests.golden_util import GOLDEN, model_path

CASES = ["qs_rsii_len3k", "err_onthq_len3k"]


def load_fixture(name):
    with open(os.path.join(GOLDEN, "stats", name + ".json")) as f:
        meta = json.load(f)
    z = np.load(os.path.join(GOLDEN, "stats", name + ".npz"))
    fix = {k: z[k] for k in z.files}
    fix["lengths"] = fix["lengths"].astype(np.int64)
    return meta, fix


def okw(meta):
    a = meta["extra_args"]
    return dict(len_mean=float(a[a.index("--length-mean") + 1]), len_sd=float(a[a.index("--length-sd") + 1]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_philox_matches_reference_distributions(name):
    meta, fix = load_fixture(name)
    o = O.Oracle(meta["method"], model_path(meta["model"]), **okw(meta))
    o.rng_philox(99)
    genome = R.synth_genome(78, [("s1", meta["genome_bp"])])[0][1]  # an independent random genome
    o.set_sequence(genome, 1)
    reads, maf, st = o.simulate_wgs(meta["depth"])
    got = SU.parse_outputs(reads, maf)
    ok, res = SU.compare(got, fix, meta["method"])
    assert ok, res
    assert got["plus"] == (got["n"] + 1) // 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_philox_matches_reference_distributions(name):
    from pbsim_b200 import capi, simulator
    meta, fix = load_fixture(name)
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params(meta["method"], **okw(meta)), model_path(meta["model"]))
    eng = simulator.Engine(0)
    eng.set_model(hm)
    eng.set_synthetic_sequence(meta["genome_bp"], 1, 4242)
    reads, maf, st, _ = eng.simulate(int(meta["depth"] * meta["genome_bp"]), rng_mode=capi.RNG_PHILOX, seed=123)
    got = SU.parse_outputs(reads, maf)
    ok, res = SU.compare(got, fix, meta["method"])
    eng.close()
    assert ok, res


# ---- the BASELINE.json configurations (fixtures: oracle/make_stats_fixtures.py, two reference seeds each) -----------
BASELINE_CASES = ["c3_qs_ont_50k", "c2_err_onthq_default", "c5_err_sequel_pass10", "c4_trans_qs_rsii"]
ALPHA = 0.002   # every KS / chi-square test; 4 configurations x 7 tests: a true-null failure in about 1 run in 18


def _split(fix, tag):
    keys = ("lengths", "accuracy", "err_accuracy", "qv_hist", "qv_thin", "events", "per_read", "plus", "n")
    out = {k: fix[tag + k] for k in keys}
    out["lengths"] = out["lengths"].astype(np.int64)
    return out


def _pass0(st, stride):
    """multi-pass runs: the passes of a read share its window and accuracy, so the read-level tests use pass 0 only"""
    if stride == 1:
        return st
    out = dict(st)
    for k in ("lengths", "accuracy", "err_accuracy", "per_read"):
        out[k] = st[k][::stride]
    out["events"] = out["per_read"].sum(0)
    return out


def _okw_baseline(meta):
    a = meta["extra"]
    kw = {}
    if "--length-mean" in a:
        kw.update(len_mean=float(a[a.index("--length-mean") + 1]), len_sd=float(a[a.index("--length-sd") + 1]),
                  len_max=int(a[a.index("--length-max") + 1]))
    if "--difference-ratio" in a:
        kw["ratio"] = tuple(int(x) for x in a[a.index("--difference-ratio") + 1].split(":"))
    if "--pass-num" in a:
        kw["pass_num"] = int(a[a.index("--pass-num") + 1])
    return kw


@pytest.mark.parametrize("name", BASELINE_CASES)
def test_reference_seed_to_seed_passes_the_same_bars(name):
    """calibration: two runs of the UNMODIFIED reference with different seeds pass the comparison the engine has to pass"""
    meta, fix = load_fixture(name)
    stride = _okw_baseline(meta).get("pass_num", 1)
    ok, res = SU.compare(_pass0(_split(fix, "b_"), stride), _pass0(_split(fix, ""), stride), meta["method"], alpha=ALPHA)
    assert ok, res


def _oracle_run(meta, seed):
    kw = _okw_baseline(meta)
    o = O.Oracle(meta["method"], model_path(meta["model"]), **kw)
    o.rng_philox(seed)
    if meta["strategy"] == "wgs":
        o.set_sequence(R.synth_genome(78, [("s1", meta["genome_bp"])])[0][1], 1)  # an independent random genome
        reads, maf, st = o.simulate_wgs(meta["depth"])
    else:
        reads, maf, st = o.simulate_set("trans", R.synth_transcripts(4242, meta["n_transcripts"], meta["n_reads"]))
    if kw.get("pass_num", 1) > 1:
        reads = SU.sam_to_fastq(reads)
    return SU.parse_outputs(reads, maf), kw.get("pass_num", 1)


@pytest.mark.parametrize("name", BASELINE_CASES)
def test_oracle_philox_matches_reference_on_baseline_configs(name):
    meta, fix = load_fixture(name)
    got, stride = _oracle_run(meta, 99)
    ok, res = SU.compare(_pass0(got, stride), _pass0(_split(fix, ""), stride), meta["method"], alpha=ALPHA)
    assert ok, res


@pytest.mark.gpu
@pytest.mark.parametrize("name", BASELINE_CASES)
def test_engine_philox_matches_reference_on_baseline_configs(name):
    from pbsim_b200 import capi, simulator
    meta, fix = load_fixture(name)
    kw = _okw_baseline(meta)
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params(meta["method"], **kw), model_path(meta["model"]))
    eng = simulator.Engine(0)
    eng.set_model(hm)
    if meta["strategy"] == "wgs":
        eng.set_synthetic_sequence(meta["genome_bp"], 1, 4242)
        reads, maf, st, _ = eng.simulate(int(meta["depth"] * meta["genome_bp"]), rng_mode=capi.RNG_PHILOX, seed=123)
    else:
        eng.set_seqset("trans", R.synth_transcripts(4242, meta["n_transcripts"], meta["n_reads"]), [0.0] + [1.0] * 10 + [0.0])
        reads, maf, st, _ = eng.simulate(0, rng_mode=capi.RNG_PHILOX, seed=123)
    eng.close()
    stride = kw.get("pass_num", 1)
    if stride > 1:
        reads = SU.sam_to_fastq(reads)
    got = SU.parse_outputs(reads, maf)
    ok, res = SU.compare(_pass0(got, stride), _pass0(_split(fix, ""), stride), meta["method"], alpha=ALPHA)
    assert ok, res


@pytest.mark.skipif(not os.path.exists(R.REF_BIN), reason="oracle/_ref/pbsim is not built here")
def test_sample_method_philox_matches_live_reference_distributions(tmp_path):
    """--method sample: PHILOX mode (oracle == engine byte for byte, tests/test_gpu_sample.py) against a run of the
    unmodified reference on the same pool and another seed — same tolerances as the qshmm fixture comparison.
    The pool is the reference's own qshmm output (as a user would make one)."""
    glen = 600000
    genome = R.synth_genome(321, [("s1", glen)])[0][1]
    fa = str(tmp_path / "g.fa")
    R.write_fasta(fa, [("s1", genome)])
    src = R.run_reference(["--strategy", "wgs", "--method", "qshmm", "--qshmm", model_path("QSHMM-RSII.model"), "--genome", fa,
                           "--depth", "2", "--seed", "7", "--length-mean", "3000", "--length-sd", "2000"])
    fq = src["files"]["out_0001.fq.gz"]
    (tmp_path / "sample.fq").write_bytes(fq)
    ref = R.run_reference(["--strategy", "wgs", "--method", "sample", "--sample", str(tmp_path / "sample.fq"), "--genome", fa,
                           "--depth", "8", "--seed", "2024"])
    assert ref["returncode"] == 0
    want = SU.parse_outputs(ref["files"]["out_0001.fq.gz"], ref["files"]["out_0001.maf.gz"])
    pool = O.sample_pool(fq)
    o = O.Oracle("sample", None)
    o.rng_philox(99)
    o.set_sequence(R.synth_genome(322, [("s1", glen)])[0][1], 1)  # an independent random genome
    reads, maf, st = o.simulate_sample(8.0, pool)
    got = SU.parse_outputs(reads, maf)
    ok, res = SU.compare(got, want, "qshmm")
    assert ok, res
    assert got["plus"] == (got["n"] + 1) // 2
