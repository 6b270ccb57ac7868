"""CPU: the C++ driver rejects what the reference rejects, with the same message and exit status, before any GPU is
touched.  tests/golden/cli_errors.json holds what the UNMODIFIED reference printed for 37 invocations (option
validation of set_sim_param pbsim.cpp:1451-1688, missing files, a reference sequence shorter than 100 bases, the
sample FASTQ filter and its statistics block on files with long lines / without a final line feed)."""
import json
import os
import re
import shutil
import subprocess

import pytest

import __graft_entry__ as G
from tests.golden_util import GOLDEN, model_path

with open(os.path.join(GOLDEN, "cli_errors.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    shutil.copy(model_path("QSHMM-RSII.model"), d / "QSHMM-RSII.model")
    shutil.copy(model_path("ERRHMM-ONT.model"), d / "ERRHMM-ONT.model")
    (d / "tiny.fa").write_text(">s\nACGTACGTAC\n")
    (d / "tiny.fq").write_text("@r1\nACGT\n+\nIIII\n@r2\nACGTA\n+\nIIIII\n")
    (d / "nonl.fq").write_text("@r1\nACGTAC\n+\nIIIIII\n@r2\nACGTA\n+\n55555")
    (d / "long.fq").write_text("@r1\n" + "ACGT" * 6250 + "\n+\n" + "5I+?" * 6250 + "\n@r2\n" + "A" * 10239 + "\n+\n"
                               + "9" * 10239 + "\n@r3\nACGT\n+\n!!!!\n")
    return d


def _norm(text, args):
    if "--seed" not in args:  # the default seed is the Unix time (pbsim.cpp:253)
        text = re.sub(r"(?m)^seed : -?\d+$", "seed : <time>", text)
    return text


@pytest.mark.parametrize("case", CASES, ids=[" ".join(c["args"][-3:]) for c in CASES])
def test_driver_rejects_like_the_reference(case, workdir):
    G.build_engine()
    exe = G.build_driver()
    p = subprocess.run([exe] + case["args"], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode == case["returncode"]
    assert _norm(p.stderr.decode(), case["args"]) == _norm(case["stderr"], case["args"])
    assert p.stdout.decode() == case["stdout"]


@pytest.mark.parametrize("name", ["tr_qs_rsii_multipass_long", "tm_qs_rsii_quirks", "tr_err_ont_hpbias"])
def test_driver_prints_the_reference_blocks_before_it_needs_a_gpu(name, tmp_path):
    """without a GPU the driver stops at engine creation — after the parameter block and the transcript / template
    statistics (get_transcript_inf :1075, get_templ_inf :1366), which must read like the reference's"""
    import gzip
    import torch
    from tests.golden_util import SetCase
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run would continue")
    c = SetCase(name)
    G.build_engine()
    exe = G.build_driver()
    with gzip.open(os.path.join(c.dir, "input.txt.gz"), "rb") as f:
        (tmp_path / "input.txt").write_bytes(f.read())
    args = [exe, "--strategy", c.strategy, "--method", c.method, "--" + c.method, c.model,
            "--transcript" if c.strategy == "trans" else "--template", "input.txt", "--seed", str(c.seed)] \
        + list(c.meta["extra_args"]) + ["--prefix", "out"]
    p = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode != 0
    got = p.stderr.decode()
    assert "no CPU fallback" in got
    head = got.split("ERROR: no usable CUDA device")[0]
    want = c.stderr.split(":::: Simulation stats ::::")[0]

    def norm(text):
        return "\n".join(("%s : <model>" % c.method) if ln.startswith(c.method + " : ") else ln for ln in text.split("\n"))
    assert norm(head) == norm(want)


@pytest.mark.parametrize("name", ["sample_basic", "sample_quirks"])
def test_sample_driver_prints_the_reference_blocks_before_it_needs_a_gpu(name, tmp_path):
    """--method sample: parameters, the sample FASTQ statistics (get_sample_inf / print_sample_stats, :1155-1358) and the
    reference statistics read like the reference's; then the driver stops for want of a GPU"""
    import gzip
    import torch
    from tests.golden_util import SampleCase
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run would continue")
    c = SampleCase(name)
    G.build_engine()
    exe = G.build_driver()
    with gzip.open(os.path.join(c.dir, "genome.fa.gz"), "rb") as f:
        (tmp_path / "genome.fa").write_bytes(f.read())
    (tmp_path / "sample.fq").write_bytes(c.sample_fastq)
    args = [exe, "--strategy", "wgs", "--method", "sample", "--sample", "sample.fq", "--genome", "genome.fa", "--depth",
            str(c.depth), "--seed", str(c.seed)] + list(c.meta["extra_args"]) + ["--prefix", "out"]
    p = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode != 0
    got = p.stderr.decode()
    assert "no CPU fallback" in got
    assert got.split("ERROR: no usable CUDA device")[0] == c.stderr.split(":::: Simulation stats (ref.1) ::::")[0]


def test_parallel_fasta_reader_equals_the_line_by_line_reader_on_a_large_file(tmp_path):
    """the driver's two FASTA readers (mapped file + a thread per record; the fgets loop that restates
    get_genome_inf, pbsim.cpp:896-991) on a 40 MB multi-FASTA with ragged line widths, blank lines, CRLF records,
    lower case and a last line without line feed: same .ref files, same report.  (Without a GPU the driver stops at
    engine creation, after the input files are read.)"""
    import subprocess
    import numpy as np
    import __graft_entry__ as G
    from tests.golden_util import model_path
    G.build_engine()
    drv = G.build_driver()
    rng = np.random.default_rng(5)
    chunks = []
    for t in range(37):
        glen = int(rng.integers(200, 2500000))
        s = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)[rng.integers(0, 10 if t % 4 == 0 else 4, glen)].tobytes()
        width = int(rng.choice([50, 60, 70, 80, 1000, 10000]))
        eol = b"\r\n" if t % 9 == 3 else b"\n"
        chunks.append(b">contig_%d some text" % t + (b" x" * 80 if t % 11 == 5 else b"") + eol)
        for i in range(0, glen, width):
            chunks.append(s[i:i + width] + eol)
            if rng.random() < 0.001:
                chunks.append(eol)
    data = b"".join(chunks)[:-1]  # the last line has no line feed
    args = ["--strategy", "wgs", "--method", "qshmm", "--qshmm", model_path("QSHMM-RSII.model"), "--genome", "g.fa",
            "--depth", "0.1", "--seed", "1", "--prefix", "out"]
    res = {}
    for who, env in (("parallel", {"PBSIM_INGEST_PARALLEL_MIN": "1"}), ("sequential", {"PBSIM_INGEST_PARALLEL_MIN": "1000000000000"})):
        d = tmp_path / who
        d.mkdir()
        (d / "g.fa").write_bytes(data)
        p = subprocess.run([drv] + args, cwd=d, env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           timeout=300)
        head = p.stderr.decode(errors="replace").split("ERROR: no usable CUDA device")[0].split(":::: Simulation stats")[0]
        refs = {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.endswith(".ref")}
        res[who] = (head, refs)
    assert len(res["parallel"][1]) == 37
    assert res["parallel"][0] == res["sequential"][0]
    assert res["parallel"][1] == res["sequential"][1]
    assert "ref.37 (len:" in res["parallel"][0]
