#!/usr/bin/env python
"""TEST INFRASTRUCTURE — distribution fixtures of the BASELINE.json configurations for the PHILOX-mode
statistical parity tests (tests/test_statistical_parity.py).

Every fixture is a reduction (tests/stats_util.py) of TWO runs of the UNMODIFIED reference (oracle/_ref/pbsim,
seeds 2024 and 2025) on scaled-down inputs with the configuration's own options:
  c3  WGS qshmm QSHMM-ONT, --length-mean 50000 --length-sd 35000 --length-max 1000000 --difference-ratio 39:24:36
  c4  --strategy trans qshmm QSHMM-RSII on a synthetic transcript table (oracle/refrun.synth_transcripts)
  c5  WGS errhmm ERRHMM-SEQUEL --pass-num 10 (SAM records)
  c2  WGS errhmm ERRHMM-ONT-HQ at the defaults (accuracy 0.85: reads below the model's range are thinned / thickened)
The second run calibrates nothing; it lets the tests show that reference-vs-reference passes the same bars.
Run where /root/reference is mounted:   python -m oracle.make_stats_fixtures
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refrun as R  # noqa: E402
from oracle.make_golden import DATA, GOLDEN, toolchain_stamp  # noqa: E402
from tests import stats_util as SU  # noqa: E402

CASES = {
    "c3_qs_ont_50k": dict(method="qshmm", model="QSHMM-ONT.model", strategy="wgs", genome_bp=3000000, depth=40,
                          extra=["--length-mean", "50000", "--length-sd", "35000", "--length-max", "1000000",
                                 "--difference-ratio", "39:24:36"]),
    "c2_err_onthq_default": dict(method="errhmm", model="ERRHMM-ONT-HQ.model", strategy="wgs", genome_bp=1500000,
                                 depth=10, extra=[]),
    "c5_err_sequel_pass10": dict(method="errhmm", model="ERRHMM-SEQUEL.model", strategy="wgs", genome_bp=1000000,
                                 depth=3, extra=["--pass-num", "10"]),
    "c4_trans_qs_rsii": dict(method="qshmm", model="QSHMM-RSII.model", strategy="trans", n_transcripts=400,
                             n_reads=12000, extra=[]),
}


def run_case(name, c, seed):
    tmp = tempfile.mkdtemp()
    if c["strategy"] == "wgs":
        fa = os.path.join(tmp, "g.fa")
        R.write_fasta(fa, R.synth_genome(77, [("s1", c["genome_bp"])]))
        args = ["--strategy", "wgs", "--method", c["method"], "--" + c["method"], os.path.join(DATA, c["model"]),
                "--genome", fa, "--depth", str(c["depth"]), "--seed", str(seed)] + c["extra"]
        res = R.run_reference(args)
        assert res["returncode"] == 0, res["stderr"]
        reads = res["files"]["out_0001.bam" if "--pass-num" in c["extra"] else "out_0001.fq.gz"]
        maf = res["files"]["out_0001.maf.gz"]
    else:
        tsv = os.path.join(tmp, "t.tsv")
        R.write_transcripts(tsv, R.synth_transcripts(4242, c["n_transcripts"], c["n_reads"]))
        args = ["--strategy", "trans", "--method", c["method"], "--" + c["method"], os.path.join(DATA, c["model"]),
                "--transcript", tsv, "--seed", str(seed)] + c["extra"]
        res = R.run_reference(args)
        assert res["returncode"] == 0, res["stderr"]
        reads, maf = res["files"]["out.fq.gz"], res["files"]["out.maf.gz"]
    if reads[:1] != b"@" or b"\t4\t*\t0\t255\t" in reads[:400]:
        reads = SU.sam_to_fastq(reads)
    return SU.parse_outputs(reads, maf)


def main():
    sdir = os.path.join(GOLDEN, "stats")
    os.makedirs(sdir, exist_ok=True)
    stamp = toolchain_stamp()
    for name, c in CASES.items():
        out = {}
        for tag, seed in (("", 2024), ("b_", 2025)):
            st = run_case(name, c, seed)
            for k, v in SU.reduce_for_fixture(st).items():
                out[tag + k] = v
            print(name, seed, st["n"], "reads", st["events"].tolist())
        np.savez_compressed(os.path.join(sdir, name + ".npz"), **out)
        meta = dict(c)
        meta.update(seeds=[2024, 2025], toolchain=stamp)
        with open(os.path.join(sdir, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
