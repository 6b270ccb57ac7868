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
