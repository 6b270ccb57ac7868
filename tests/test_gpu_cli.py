"""GPU test of the C++ host driver (pbsim_b200/bin/pbsim): same command line as the reference, outputs
compared byte for byte with the reference's golden run (replay mode) and with the oracle (philox mode)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import oracle as O
from tests import bam_util as B
from tests.golden_util import Case

pytestmark = pytest.mark.gpu


def _run_cli(c, tmp_path, extra):
    exe = G.build_driver()
    fa = tmp_path / "genome.fa"
    with gzip.open(os.path.join(c.dir, "genome.fa.gz"), "rb") as f:
        fa.write_bytes(f.read())
    args = [exe, "--strategy", "wgs", "--method", c.method, "--" + c.method, c.model, "--genome", "genome.fa",
            "--depth", str(c.depth), "--seed", str(c.seed)] + list(c.meta["extra_args"]) + ["--prefix", "out"] + extra
    p = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    return p.stderr.decode()


@pytest.mark.parametrize("name", ["qs_rsii_quirks", "qs_ont_hpbias", "err_sequel_multipass"])
def test_cli_replay_reproduces_reference_files(name, tmp_path):
    c = Case(name)
    O.glibc_rand(c.seed, c.ndraws).tofile(tmp_path / "draws.bin")
    c.marks.astype(np.int64).tofile(tmp_path / "marks.bin")
    stderr = _run_cli(c, tmp_path, ["--rng", "replay", "--replay-draws", "draws.bin", "--replay-marks", "marks.bin"])
    for i in range(1, len(c.contigs) + 1):
        want = gzip.open(os.path.join(c.dir, "seq%d.reads.gz" % i), "rb").read()  # SAM: header included
        if c.pass_num == 1:
            got = gzip.open(tmp_path / ("out_%04d.fq.gz" % i), "rb").read()
        else:  # <prefix>_NNNN.bam like the reference (:715): decoded by the tests' own BAM reader
            text, recs = B.parse_bam((tmp_path / ("out_%04d.bam" % i)).read_bytes())
            got = text + recs
        assert got == want, "reads file differs, seq %d" % i
        assert gzip.open(tmp_path / ("out_%04d.maf.gz" % i), "rb").read() == c.maf(i)
        ref = (tmp_path / ("out_%04d.ref" % i)).read_bytes()
        assert ref == gzip.open(os.path.join(c.dir, "seq%d.ref.gz" % i), "rb").read()
    # stderr: identical up to the "System utilization" block (timings) and the engine's own trailer
    def norm(text):  # the echoed model path differs (tests unpack the model to a temp dir)
        head = text.split(":::: System utilization ::::")[0]
        return "\n".join(("%s : <model>" % c.method) if ln.startswith(c.method + " : ") else ln
                         for ln in head.split("\n"))
    assert norm(stderr) == norm(c.stderr)


def test_cli_philox_equals_oracle(tmp_path):
    c = Case("qs_rsii_basic")
    out, _ = c.run_oracle("philox")
    _run_cli(c, tmp_path, [])
    for i, o in enumerate(out, start=1):
        assert gzip.open(tmp_path / ("out_%04d.fq.gz" % i), "rb").read() == o["reads"]
        assert gzip.open(tmp_path / ("out_%04d.maf.gz" % i), "rb").read() == o["maf"]


@pytest.mark.parametrize("name", ["tr_qs_ont_hpbias", "tr_err_ont_hpbias", "tm_err_sequel_multipass",
                                  "tr_qs_rsii_multipass_long", "tm_qs_rsii_hpbias"])
def test_cli_replay_reproduces_reference_transcript_and_template_files(name, tmp_path):
    """--strategy trans / templ: the driver parses the transcript table / template FASTA itself (long lines,
    lower case, multi-line FASTA) and must write the reference's <prefix>.fq.gz|.bam text and <prefix>.maf.gz"""
    from tests.golden_util import SetCase
    c = SetCase(name)
    exe = G.build_driver()
    with gzip.open(os.path.join(c.dir, "input.txt.gz"), "rb") as f:
        (tmp_path / "input.txt").write_bytes(f.read())
    O.glibc_rand(c.seed, c.ndraws).tofile(tmp_path / "draws.bin")
    c.marks.astype(np.int64).tofile(tmp_path / "marks.bin")
    args = [exe, "--strategy", c.strategy, "--method", c.method, "--" + c.method, c.model,
            "--transcript" if c.strategy == "trans" else "--template", "input.txt", "--seed", str(c.seed)] \
        + list(c.meta["extra_args"]) + ["--prefix", "out", "--rng", "replay", "--replay-draws", "draws.bin",
                                        "--replay-marks", "marks.bin"]
    p = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    want = gzip.open(os.path.join(c.dir, "reads.gz"), "rb").read()  # SAM: header included
    if c.pass_num == 1:
        got = gzip.open(tmp_path / "out.fq.gz", "rb").read()
    else:
        text, recs = B.parse_bam((tmp_path / "out.bam").read_bytes())
        got = text + recs
        if name == "tm_err_sequel_multipass":  # BAM has no lower-case bases (the first base of a template may be)
            def up(sam):
                rows = []
                for ln in sam.decode().splitlines():
                    f = ln.split("\t")
                    if len(f) > 9:
                        f[9] = f[9].upper()
                    rows.append("\t".join(f))
                return rows
            got, want = up(got), up(want)
    assert got == want
    assert gzip.open(tmp_path / "out.maf.gz", "rb").read() == c.maf()

    def norm(text):
        head = text.split(":::: System utilization ::::")[0]
        return "\n".join(("%s : <model>" % c.method) if ln.startswith(c.method + " : ") else ln
                         for ln in head.split("\n"))
    assert norm(p.stderr.decode()) == norm(c.stderr)


def test_cli_host_gzip_writes_sam_text_for_multipass(tmp_path):
    """--gzip host: zlib threads on the host, multi-pass records as SAM text in <prefix>_NNNN.sam.gz"""
    c = Case("err_sequel_multipass")
    O.glibc_rand(c.seed, c.ndraws).tofile(tmp_path / "draws.bin")
    c.marks.astype(np.int64).tofile(tmp_path / "marks.bin")
    _run_cli(c, tmp_path, ["--rng", "replay", "--replay-draws", "draws.bin", "--replay-marks", "marks.bin",
                           "--gzip", "host"])
    for i in range(1, len(c.contigs) + 1):
        got = gzip.open(tmp_path / ("out_%04d.sam.gz" % i), "rb").read()
        assert got == gzip.open(os.path.join(c.dir, "seq%d.reads.gz" % i), "rb").read()
        assert gzip.open(tmp_path / ("out_%04d.maf.gz" % i), "rb").read() == c.maf(i)


def test_cli_two_processes_shard_the_sequences(tmp_path):
    """--rank / --world: one process per GPU, sequences dealt round-robin; the union of the two processes' files
    equals the files of a single process (here both run on the one GPU)"""
    c = Case("qs_rsii_quirks")  # three contigs, homopolymers >= 11 (the aliased bias cell runs across sequences)
    out, _ = c.run_oracle("philox")
    for rank in (0, 1):
        _run_cli(c, tmp_path, ["--rank", str(rank), "--world", "2"])
    for i, o in enumerate(out, start=1):
        assert gzip.open(tmp_path / ("out_%04d.fq.gz" % i), "rb").read() == o["reads"]
        assert gzip.open(tmp_path / ("out_%04d.maf.gz" % i), "rb").read() == o["maf"]


@pytest.mark.parametrize("name", ["sample_basic", "sample_quirks"])
def test_cli_sample_method_reproduces_reference_files(name, tmp_path):
    """--method sample: the driver filters the FASTQ (get_sample_inf), the engine copies the pool; files and stderr are
    the reference's.  Then the profile round trip: --sample + --sample-profile-id stores the filtered reads,
    --sample-profile-id alone reuses them (pbsim.cpp:1565-1616) and must simulate the same reads."""
    from tests.golden_util import SampleCase
    c = SampleCase(name)
    exe = G.build_driver()
    with gzip.open(os.path.join(c.dir, "genome.fa.gz"), "rb") as f:
        (tmp_path / "genome.fa").write_bytes(f.read())
    (tmp_path / "sample.fq").write_bytes(c.sample_fastq)
    O.glibc_rand(c.seed, c.ndraws).tofile(tmp_path / "draws.bin")
    c.marks.astype(np.int64).tofile(tmp_path / "marks.bin")
    common = ["--strategy", "wgs", "--method", "sample", "--genome", "genome.fa", "--depth", str(c.depth), "--seed",
              str(c.seed)] + list(c.meta["extra_args"])
    replay = ["--rng", "replay", "--replay-draws", "draws.bin", "--replay-marks", "marks.bin"]

    def run(extra, prefix):
        p = subprocess.run([exe] + common + extra + ["--prefix", prefix], cwd=tmp_path, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=600)
        assert p.returncode == 0, p.stderr.decode()
        return p.stderr.decode()

    stderr = run(["--sample", "sample.fq"] + replay, "out")
    for i in range(1, len(c.contigs) + 1):
        assert gzip.open(tmp_path / ("out_%04d.fq.gz" % i), "rb").read() == c.reads(i), "reads file differs, seq %d" % i
        assert gzip.open(tmp_path / ("out_%04d.maf.gz" % i), "rb").read() == c.maf(i)
    assert stderr.split(":::: System utilization ::::")[0] == c.stderr.split(":::: System utilization ::::")[0]
    # store, then reuse
    run(["--sample", "sample.fq", "--sample-profile-id", "p1"] + replay, "st")
    stats = (tmp_path / "sample_profile_p1.stats").read_text()
    assert stats.startswith("num\t%d\nlen_total\t%d\n" % (len(c.pool), sum(len(q) for q in c.pool)))
    assert (tmp_path / "sample_profile_p1.fastq").read_bytes() == b"".join(q + b"\n" for q in c.pool)
    err2 = run(["--sample-profile-id", "p1"] + replay, "re")
    assert "file name : sample_profile_p1.fastq\n\n:: filtered reads ::\nread num. : %d\n" % len(c.pool) in err2
    for i in range(1, len(c.contigs) + 1):
        for ext in ("fq.gz", "maf.gz"):
            assert gzip.open(tmp_path / ("st_%04d.%s" % (i, ext)), "rb").read() == \
                gzip.open(tmp_path / ("out_%04d.%s" % (i, ext)), "rb").read()
            assert gzip.open(tmp_path / ("re_%04d.%s" % (i, ext)), "rb").read() == \
                gzip.open(tmp_path / ("out_%04d.%s" % (i, ext)), "rb").read()
    # storing over an existing profile is refused like the reference does
    p = subprocess.run([exe] + common + ["--sample", "sample.fq", "--sample-profile-id", "p1"], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert p.returncode != 0 and "ERROR: sample_profile_p1.fastq exists." in p.stderr.decode()
