"""CPU: the oracle against the UNMODIFIED reference run live (oracle/_ref/pbsim, built from /root/reference by
`make -C oracle ref`; skipped where that binary is absent).  Seeded random configurations — the ones the engine is
fuzzed with in tests/test_fuzz_core_cpu.py and tests/test_gpu_zz_fuzz.py — go through the reference's own command line;
FASTQ / SAM, MAF and the statistics block of the oracle (glibc rand() restatement, same seed) must equal the
reference's byte for byte.  This extends the pin of the oracle beyond the committed golden runs."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refrun as R
from tests.golden_util import model_path
from tests.test_fuzz_core_cpu import _config

pytestmark = pytest.mark.skipif(not os.path.exists(R.REF_BIN), reason="oracle/_ref/pbsim is not built here")


@pytest.mark.parametrize("k", range(48))
def test_oracle_equals_live_reference(k, tmp_path):
    cfg = _config(k)
    okw, genome, method = cfg["okw"], cfg["genome"], cfg["method"]
    fa = str(tmp_path / "genome.fa")
    R.write_fasta(fa, [("g", genome)])
    args = ["--strategy", "wgs", "--method", method, "--genome", fa, "--depth", repr(cfg["depth"]), "--seed", str(cfg["seed"]),
            "--length-min", str(okw["len_min"]), "--length-max", str(okw["len_max"]),
            "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"])]
    pool = None
    if method == "sample":
        lens = cfg["rng"].integers(100, 3000, int(cfg["rng"].integers(4, 40)))
        quals = [bytes(cfg["rng"].integers(33 + 3, 33 + 25, int(n)).astype(np.uint8)) for n in lens]
        fq = b"".join(b"@r%d\n" % i + b"A" * len(q) + b"\n+\n" + q + b"\n" for i, q in enumerate(quals))
        (tmp_path / "sample.fq").write_bytes(fq)
        args += ["--sample", str(tmp_path / "sample.fq"), "--accuracy-min", "0", "--accuracy-max", "1"]
        pool = O.sample_pool(fq, len_min=okw["len_min"], len_max=okw["len_max"], accuracy_min=0.0, accuracy_max=1.0)
        assert len(pool) >= 2
    else:
        args += ["--" + method, model_path(cfg["model"]), "--length-mean", repr(okw["len_mean"]), "--length-sd",
                 repr(okw["len_sd"]), "--accuracy-mean", "%.2f" % okw["accuracy_mean"], "--pass-num", str(okw["pass_num"])]
    ref = R.run_reference(args)
    try:
        o = O.Oracle(method, model_path(cfg["model"]) if cfg["model"] else None, **okw)
        o.rng_glibc(cfg["seed"])
        if okw["hp_del_bias"] != 1.0:
            o.hp_bias_prepass([genome])
        o.set_sequence(genome, 1)
        reads, maf, st = o.simulate_sample(cfg["depth"], pool) if pool is not None else o.simulate_wgs(cfg["depth"])
    except RuntimeError as e:
        # what the oracle refuses the reference must refuse too
        assert ref["returncode"] != 0, "oracle failed (%s) where the reference ran" % e
        return
    assert ref["returncode"] == 0, ref["stderr"][-400:]
    multi = okw.get("pass_num", 1) > 1
    got = ref["files"]["out_0001.bam" if multi else "out_0001.fq.gz"]
    if multi:  # the two header lines main() writes (:721-722)
        got = got[got.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n"):]
    assert reads == got, "reads differ from the live reference"
    assert maf == ref["files"]["out_0001.maf.gz"], "MAF differs from the live reference"
    assert O.format_stats(st, 1) == R.split_stats_blocks(ref["stderr"])[1]


@pytest.mark.parametrize("k", range(24))
def test_engine_core_replays_live_reference(k, tmp_path):
    """the engine's per-read core (tests/hostsim) fed the rand() stream of a live reference run (the -include hook of
    oracle/ref_hooks logs every draw) reproduces that run's files"""
    from pbsim_b200 import capi
    from tests import hostsim_util as H
    cfg = _config(k)
    okw, genome, method = cfg["okw"], cfg["genome"], cfg["method"]
    if not os.path.exists(R.REF_BIN_LOG):
        pytest.skip("oracle/_ref/pbsim_logrand is not built here")
    fa = str(tmp_path / "genome.fa")
    R.write_fasta(fa, [("g", genome)])
    args = ["--strategy", "wgs", "--method", method, "--genome", fa, "--depth", repr(cfg["depth"]), "--seed", str(cfg["seed"]),
            "--length-min", str(okw["len_min"]), "--length-max", str(okw["len_max"]),
            "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"])]
    pool = None
    if method == "sample":
        lens = cfg["rng"].integers(100, 3000, int(cfg["rng"].integers(4, 40)))
        quals = [bytes(cfg["rng"].integers(33 + 3, 33 + 25, int(n)).astype(np.uint8)) for n in lens]
        fq = b"".join(b"@r%d\n" % i + b"A" * len(q) + b"\n+\n" + q + b"\n" for i, q in enumerate(quals))
        (tmp_path / "sample.fq").write_bytes(fq)
        args += ["--sample", str(tmp_path / "sample.fq"), "--accuracy-min", "0", "--accuracy-max", "1"]
        pool, _ = capi.sample_filter(H.lib(), fq, len_min=okw["len_min"], len_max=okw["len_max"], accuracy_min=0.0,
                                     accuracy_max=1.0)
    else:
        args += ["--" + method, model_path(cfg["model"]), "--length-mean", repr(okw["len_mean"]), "--length-sd",
                 repr(okw["len_sd"]), "--accuracy-mean", "%.2f" % okw["accuracy_mean"], "--pass-num", str(okw["pass_num"])]
    ref = R.run_reference(args, logrand=True)
    if ref["returncode"] != 0:
        pytest.skip("the reference rejects this configuration")
    try:
        hm = capi.HostModel(H.lib(), capi.host_params(method, **okw), model_path(cfg["model"]) if cfg["model"] else None)
    except RuntimeError as e:
        pytest.fail("the product's table builder fails (%s) where the reference ran" % e)
    o = O.Oracle(method, model_path(cfg["model"]) if cfg["model"] else None, **okw)  # ingest only: upper case, hp, bias
    if okw["hp_del_bias"] != 1.0:
        o.hp_bias_prepass([genome])
    o.set_sequence(genome, 1)
    sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_REPLAY, 0, ref["draws"], int(cfg["depth"] * len(genome)),
                pool=pool, batch_reads=int(cfg["rng"].integers(1, 20)))
    reads, maf = H.records_from_events(hm, sub, o.seq_upper(), 1)
    multi = okw.get("pass_num", 1) > 1
    got = ref["files"]["out_0001.bam" if multi else "out_0001.fq.gz"]
    if multi:
        got = got[got.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n"):]
    assert reads == got
    assert maf == ref["files"]["out_0001.maf.gz"]
    assert [s["draw_start"] for s in sub] == [0] + [int(x) for x in ref["marks"][:len(sub) - 1]]


@pytest.mark.parametrize("k", range(24))
def test_oracle_equals_live_reference_on_sequence_sets(k, tmp_path):
    """--strategy trans / templ: random transcript tables (expression counts on both strands, lower-case sequences,
    IUPAC codes, homopolymers, lines longer than an fgets buffer) and template FASTA files"""
    rng = np.random.default_rng(7700 + k)
    strategy = ["trans", "templ"][k % 2]
    method = ["qshmm", "errhmm"][(k // 2) % 2]
    model = str(rng.choice(["QSHMM-RSII.model", "QSHMM-ONT.model", "QSHMM-ONT-HQ.model"] if method == "qshmm" else
                           ["ERRHMM-RSII.model", "ERRHMM-ONT.model", "ERRHMM-ONT-HQ.model", "ERRHMM-SEQUEL.model"]))
    mean = float(rng.integers(800, 4000))
    okw = dict(len_min=100, len_max=int(rng.integers(3000, 30000)), ratio=tuple(int(x) for x in rng.integers(1, 60, 3)),
               hp_del_bias=float(rng.choice([1.0, 1.0, 3.0])), len_mean=mean, len_sd=float(rng.uniform(0.3, 0.9)) * mean,
               pass_num=int(rng.choice([1, 1, 3])), accuracy_mean=float(rng.integers(82, 97)) / 100.0, accuracy_mean_set=True)
    seqset = R.synth_set(300 + k, int(rng.integers(5, 30)), len_lo=150, len_hi=int(rng.integers(600, 5000)),
                         max_exp=int(rng.integers(1, 5)), lowercase_first=float(rng.choice([0.0, 0.3])),
                         iupac=float(rng.choice([0.0, 0.002])), hp_plants=int(rng.integers(0, 4)),
                         long_every=int(rng.choice([0, 7])))
    if strategy == "templ":
        seqset = [(n, 1, 0, s) for n, _, _, s in seqset]
        # a template longer than --length-max overruns the reference's read buffers (malloc(len_max * 2 + 1), and half
        # of that for mut.hp, :5488-5531): undefined behaviour there, so the comparison stays inside the buffers
        okw["len_max"] = max(okw["len_max"], max(len(x[3]) for x in seqset) + 1)
    seed = int(rng.integers(1, 1 << 30))
    inp = str(tmp_path / "input.txt")
    (R.write_transcripts if strategy == "trans" else R.write_templates)(inp, seqset)
    args = ["--strategy", strategy, "--method", method, "--" + method, model_path(model),
            "--transcript" if strategy == "trans" else "--template", inp, "--seed", str(seed),
            "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"]),
            "--accuracy-mean", "%.2f" % okw["accuracy_mean"], "--pass-num", str(okw["pass_num"])]
    # (templates are read in full whatever --length-max says, but the SD of the statistics block only counts
    # lengths up to it, :3568-3575)
    args += ["--length-min", str(okw["len_min"]), "--length-max", str(okw["len_max"])]
    if strategy == "trans":
        args += ["--length-mean", repr(okw["len_mean"]), "--length-sd", repr(okw["len_sd"])]
    ref = R.run_reference(args)
    try:
        o = O.Oracle(method, model_path(model), **okw)
        o.rng_glibc(seed)
        reads, maf, st = o.simulate_set(strategy, seqset)
    except RuntimeError as e:
        assert ref["returncode"] != 0, "oracle failed (%s) where the reference ran" % e
        return
    assert ref["returncode"] == 0, ref["stderr"][-400:]
    multi = okw["pass_num"] > 1
    got = ref["files"]["out.bam" if multi else "out.fq.gz"]
    if multi:
        got = got[got.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n"):]
    assert reads == got, "reads differ from the live reference"
    assert maf == ref["files"]["out.maf.gz"], "MAF differs from the live reference"
    assert O.format_stats_set(st) == R.set_stats_block(ref["stderr"])


def _config_wide(k):
    """wider corners than _config (which the GPU fuzz shares): accuracies 0.70 .. 1.00 (below / above the models' range,
    accuracy-100 reads), fixed-length reads, genomes shorter than the reads, zero entries in the ratio, up to 4 passes"""
    rng = np.random.default_rng(9100 + k)
    method = ["qshmm", "errhmm"][k % 2]
    model = str(rng.choice(["QSHMM-RSII.model", "QSHMM-ONT.model", "QSHMM-ONT-HQ.model"] if method == "qshmm" else
                           ["ERRHMM-RSII.model", "ERRHMM-ONT.model", "ERRHMM-ONT-HQ.model", "ERRHMM-SEQUEL.model"]))
    mean = float(rng.integers(300, 8000))
    ratio = [int(x) for x in rng.integers(0, 60, 3)]
    if k % 5 == 0:
        ratio[int(rng.integers(0, 3))] = 0
    if sum(ratio) == 0:
        ratio[0] = 1
    okw = dict(len_min=int(rng.choice([100, 100, 250])), len_max=int(rng.integers(3000, 40000)), ratio=tuple(ratio),
               hp_del_bias=float(rng.choice([1.0, 1.0, 1.5, 6.0])), len_mean=mean,
               len_sd=0.0 if k % 6 == 0 else float(rng.uniform(0.2, 0.95)) * mean, pass_num=int(rng.choice([1, 1, 2, 4])),
               accuracy_mean=float(rng.integers(70, 101)) / 100.0, accuracy_mean_set=True)
    glen = int(rng.integers(400, 3000)) if k % 4 == 0 else int(rng.integers(8000, 50000))
    genome = R.synth_genome(500 + k, [("g", glen)], n_runs=int(rng.integers(0, 4)), hp_plants=int(rng.integers(0, 40)),
                            lowercase_frac=float(rng.choice([0.0, 0.2])), iupac=int(rng.integers(0, 6)),
                            long_runs=(12, 30) if k % 3 == 0 else ())[0][1]
    return dict(method=method, model=model, okw=okw, genome=genome, depth=float(rng.uniform(0.5, 6.0)),
                seed=int(rng.integers(1, 1 << 30)), rng=rng)


@pytest.mark.parametrize("k", range(40))
def test_wide_corners_reference_oracle_and_engine_core_agree(k, tmp_path):
    """reference (live) == oracle (glibc) == engine core (replay of the logged draws); engine core (PHILOX) == oracle
    (PHILOX) — on the wide corner configurations; what the reference rejects, the product's table builder rejects"""
    from pbsim_b200 import capi
    from tests import hostsim_util as H
    cfg = _config_wide(k)
    okw, genome, method = cfg["okw"], cfg["genome"], cfg["method"]
    fa = str(tmp_path / "genome.fa")
    R.write_fasta(fa, [("g", genome)])
    args = ["--strategy", "wgs", "--method", method, "--genome", fa, "--depth", repr(cfg["depth"]), "--seed", str(cfg["seed"]),
            "--length-min", str(okw["len_min"]), "--length-max", str(okw["len_max"]),
            "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"]),
            "--" + method, model_path(cfg["model"]), "--length-mean", repr(okw["len_mean"]), "--length-sd",
            repr(okw["len_sd"]), "--accuracy-mean", "%.2f" % okw["accuracy_mean"], "--pass-num", str(okw["pass_num"])]
    ref = R.run_reference(args, logrand=True)
    mp = model_path(cfg["model"])
    if ref["returncode"] != 0:
        with pytest.raises(RuntimeError):
            o = O.Oracle(method, mp, **okw)
            o.rng_glibc(cfg["seed"])
            o.set_sequence(genome, 1)
            o.simulate_wgs(cfg["depth"])
        return
    o = O.Oracle(method, mp, **okw)
    o.rng_glibc(cfg["seed"])
    if okw["hp_del_bias"] != 1.0:
        o.hp_bias_prepass([genome])
    o.set_sequence(genome, 1)
    reads, maf, st = o.simulate_wgs(cfg["depth"])
    multi = okw["pass_num"] > 1
    got = ref["files"]["out_0001.bam" if multi else "out_0001.fq.gz"]
    if multi:
        got = got[got.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n"):]
    assert reads == got, "oracle reads differ from the live reference"
    assert maf == ref["files"]["out_0001.maf.gz"]
    assert O.format_stats(st, 1) == R.split_stats_blocks(ref["stderr"])[1]
    hm = capi.HostModel(H.lib(), capi.host_params(method, **okw), mp)
    quota = int(cfg["depth"] * len(genome))
    sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_REPLAY, 0, ref["draws"], quota)
    r2, m2 = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert r2 == got and m2 == ref["files"]["out_0001.maf.gz"], "engine core replay differs from the live reference"
    # PHILOX: engine core (segments forced on) against the oracle
    o.rng_philox(cfg["seed"])
    preads, pmaf, _ = o.simulate_wgs(cfg["depth"])
    L = H.lib()
    L.hostsim_use_segments(1, 1025)
    L.hostsim_set_chain_chunk(int(cfg["rng"].choice([0, 1, 3, 32])))
    try:
        sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_PHILOX, cfg["seed"], None, quota)
    finally:
        L.hostsim_use_segments(0, 2048)
        L.hostsim_set_chain_chunk(0)
    r3, m3 = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert r3 == preads and m3 == pmaf, "engine core (PHILOX) differs from the oracle"


@pytest.mark.parametrize("k", range(10))
def test_driver_parses_fasta_like_the_live_reference(k, tmp_path):
    """get_genome_inf (:896-995) through the C++ driver: random multi-FASTA files (many contigs, ragged line widths,
    lines longer than an fgets buffer, blank lines, lower case, names with spaces and longer than 128 characters) —
    the <prefix>_NNNN.ref files and everything printed before the first simulation must equal the reference's.  (The
    driver stops at engine creation on a box without a GPU; with one it runs on, the comparison is the same.)"""
    import subprocess
    import __graft_entry__ as G
    rng = np.random.default_rng(3300 + k)
    n = int(rng.integers(1, 12))
    lines = []
    for t in range(n):
        name = "ctg%d" % t + (" some description" if t % 3 == 0 else "") + ("x" * 150 if t % 5 == 4 else "")
        glen = int(rng.integers(100, 30000))
        s = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)[rng.integers(0, 10 if t % 2 else 4, glen)].tobytes()
        lines.append(b">" + name.encode())
        width = int(rng.choice([60, 70, 11000, 25000]))
        for i in range(0, glen, width):
            lines.append(s[i:i + width])
            if rng.random() < 0.05:
                lines.append(b"")
    fa = tmp_path / "genome.fa"
    fa.write_bytes(b"\n".join(lines) + (b"\n" if k % 2 == 0 else b""))
    args = ["--strategy", "wgs", "--method", "qshmm", "--qshmm", model_path("QSHMM-RSII.model"), "--genome", "genome.fa",
            "--depth", "0.5", "--seed", "3", "--length-mean", "500", "--length-sd", "200", "--length-min", "100"]
    rdir, ddir = tmp_path / "r", tmp_path / "d"
    rdir.mkdir()
    ddir.mkdir()
    for d in (rdir, ddir):
        (d / "genome.fa").write_bytes(fa.read_bytes())
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    pr = subprocess.run([R.REF_BIN] + args + ["--prefix", "out"], cwd=rdir, env=env, stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, timeout=300)
    G.build_engine()
    # even k: the parallel reader (mapped file, a thread per record) is forced onto these small files; it declines
    # files with lines longer than an fgets buffer and leaves them to the line-by-line reader
    pd = subprocess.run([G.build_driver()] + args + ["--prefix", "out"], cwd=ddir, stdout=subprocess.PIPE,
                        env=dict(os.environ, PBSIM_INGEST_PARALLEL_MIN="1") if k % 2 == 0 else dict(os.environ),
                        stderr=subprocess.PIPE, timeout=300)
    want, got = pr.stderr.decode(), pd.stderr.decode()
    cut = ":::: Simulation stats (ref.1) ::::"
    head_w = want.split(cut)[0]
    head_g = got.split("ERROR: no usable CUDA device")[0].split(cut)[0]
    assert head_g == head_w
    if pr.returncode != 0:  # e.g. a contig shorter than 100 bases: same message, same status
        assert pd.returncode == pr.returncode
        return
    refs = sorted(f for f in os.listdir(rdir) if f.endswith(".ref"))
    assert refs and refs == sorted(f for f in os.listdir(ddir) if f.endswith(".ref"))
    for f in refs:
        assert (ddir / f).read_bytes() == (rdir / f).read_bytes(), f


@pytest.mark.parametrize("k", range(8))
def test_driver_parses_sets_like_the_live_reference(k, tmp_path):
    """get_transcript_inf (:1075) / get_templ_inf (:1366) through the C++ driver on random tables / FASTA files: the
    statistics blocks printed before the simulation equal the reference's"""
    import subprocess
    import __graft_entry__ as G
    rng = np.random.default_rng(5500 + k)
    strategy = ["trans", "templ"][k % 2]
    seqset = R.synth_set(800 + k, int(rng.integers(3, 40)), len_lo=120, len_hi=int(rng.integers(500, 6000)),
                         max_exp=int(rng.integers(1, 9)), lowercase_first=0.3, iupac=0.001, hp_plants=2,
                         long_every=int(rng.choice([0, 5])))
    inp = tmp_path / "input.txt"
    if strategy == "trans":
        R.write_transcripts(str(inp), seqset)
    else:
        R.write_templates(str(inp), seqset, width=int(rng.choice([60, 70, 12000])))
    args = ["--strategy", strategy, "--method", "errhmm", "--errhmm", model_path("ERRHMM-ONT.model"),
            "--transcript" if strategy == "trans" else "--template", "input.txt", "--seed", "5", "--length-mean", "600",
            "--length-sd", "300"]
    rdir, ddir = tmp_path / "r", tmp_path / "d"
    for d in (rdir, ddir):
        d.mkdir()
        (d / "input.txt").write_bytes(inp.read_bytes())
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    pr = subprocess.run([R.REF_BIN] + args + ["--prefix", "out"], cwd=rdir, env=env, stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, timeout=300)
    G.build_engine()
    pd = subprocess.run([G.build_driver()] + args + ["--prefix", "out"], cwd=ddir, stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, timeout=300)
    cut = ":::: Simulation stats ::::"
    assert pd.stderr.decode().split("ERROR: no usable CUDA device")[0].split(cut)[0] == pr.stderr.decode().split(cut)[0]


@pytest.mark.parametrize("k", range(20))
def test_sample_method_live_with_filters(k, tmp_path):
    """--method sample with the length and accuracy filters cutting into the FASTQ, sequences shorter than the sampled
    reads, deletion-rich and insertion-rich mixes: product filter == reference's statistics block, oracle == reference
    files, engine core (schedule + per-read core, replay of the logged draws) == reference files"""
    from pbsim_b200 import capi
    from tests import hostsim_util as H
    rng = np.random.default_rng(6600 + k)
    glen = int(rng.integers(300, 2500)) if k % 3 == 0 else int(rng.integers(5000, 40000))
    genome = R.synth_genome(40 + k, [("g", glen)], n_runs=int(rng.integers(0, 3)), hp_plants=int(rng.integers(0, 30)),
                            iupac=int(rng.integers(0, 4)), lowercase_frac=float(rng.choice([0.0, 0.15])))[0][1]
    okw = dict(len_min=int(rng.integers(100, 600)), len_max=int(rng.integers(1500, 4000)),
               ratio=tuple(int(x) for x in rng.integers(1, 60, 3)), hp_del_bias=float(rng.choice([1.0, 1.0, 2.0, 5.0])))
    amin, amax = sorted(float(x) / 100 for x in rng.integers(70, 100, 2))
    if amin == amax:
        amax = min(1.0, amax + 0.05)
    lens = rng.integers(50, 5000, int(rng.integers(10, 80)))
    recs = []
    for i, n in enumerate(lens):
        lo = int(rng.integers(2, 15))
        q = bytes(rng.integers(33 + lo, 33 + lo + int(rng.integers(2, 25)), int(n)).astype(np.uint8))
        recs.append(b"@r%d\n" % i + b"C" * int(n) + b"\n+\n" + q + b"\n")
    fq = b"".join(recs)
    (tmp_path / "sample.fq").write_bytes(fq)
    fa = str(tmp_path / "genome.fa")
    R.write_fasta(fa, [("g", genome)])
    depth = float(rng.uniform(0.5, 12.0))
    seed = int(rng.integers(1, 1 << 30))
    args = ["--strategy", "wgs", "--method", "sample", "--sample", str(tmp_path / "sample.fq"), "--genome", fa, "--depth",
            repr(depth), "--seed", str(seed), "--length-min", str(okw["len_min"]), "--length-max", str(okw["len_max"]),
            "--accuracy-min", "%.2f" % amin, "--accuracy-max", "%.2f" % amax,
            "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"])]
    ref = R.run_reference(args, logrand=True)
    tr = lambda x: int(float("%.2f" % x) * 100) * 0.01  # noqa: E731  (set_sim_param truncates to two decimals, :1620-1629)
    try:
        pool, ss = capi.sample_filter(H.lib(), fq, len_min=okw["len_min"], len_max=okw["len_max"], accuracy_min=tr(amin),
                                      accuracy_max=tr(amax))
    except RuntimeError as e:
        assert ref["returncode"] != 0 and str(e) in ref["stderr"]
        return
    assert capi.format_sample_stats(ss, str(tmp_path / "sample.fq")) in ref["stderr"]
    assert pool == O.sample_pool(fq, len_min=okw["len_min"], len_max=okw["len_max"], accuracy_min=tr(amin),
                                 accuracy_max=tr(amax))
    if len(pool) < 2:
        return  # the reference divides by zero (:1723); the engine refuses the pool
    assert ref["returncode"] == 0, ref["stderr"][-300:]
    o = O.Oracle("sample", None, **okw)
    o.rng_glibc(seed)
    if okw["hp_del_bias"] != 1.0:
        o.hp_bias_prepass([genome])
    o.set_sequence(genome, 1)
    reads, maf, st = o.simulate_sample(depth, pool)
    assert reads == ref["files"]["out_0001.fq.gz"] and maf == ref["files"]["out_0001.maf.gz"]
    assert O.format_stats(st, 1) == R.split_stats_blocks(ref["stderr"])[1]
    hm = capi.HostModel(H.lib(), capi.host_params("sample", **okw), None)
    sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_REPLAY, 0, ref["draws"], int(depth * len(genome)),
                pool=pool, batch_reads=int(rng.integers(1, 30)))
    r2, m2 = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert r2 == reads and m2 == maf


@pytest.mark.parametrize("k", range(10))
def test_multi_contig_genomes_live(k, tmp_path):
    """several sequences per run: the rand() stream, the --hp-del-bias prepass over all sequences and the hpfreq[11]
    cell that aliases hp_del_bias[0] carry over from one sequence to the next (main :673-754)"""
    from pbsim_b200 import capi
    from tests import hostsim_util as H
    rng = np.random.default_rng(8800 + k)
    method = ["qshmm", "errhmm"][k % 2]
    model = str(rng.choice(["QSHMM-RSII.model", "QSHMM-ONT.model"] if method == "qshmm" else ["ERRHMM-ONT.model", "ERRHMM-SEQUEL.model"]))
    contigs = R.synth_genome(70 + k, [("c%d" % t, int(rng.integers(300, 20000))) for t in range(int(rng.integers(2, 6)))],
                             n_runs=2, hp_plants=25, iupac=2, long_runs=(11, 12, 19, 40))
    okw = dict(len_min=100, len_max=20000, ratio=tuple(int(x) for x in rng.integers(1, 60, 3)),
               hp_del_bias=float(rng.choice([1.0, 2.0, 7.5])), len_mean=float(rng.integers(500, 3000)), len_sd=400.0,
               pass_num=int(rng.choice([1, 2])), accuracy_mean=float(rng.integers(80, 99)) / 100.0, accuracy_mean_set=True)
    depth, seed = float(rng.uniform(1.0, 4.0)), int(rng.integers(1, 1 << 30))
    fa = str(tmp_path / "genome.fa")
    R.write_fasta(fa, contigs)
    args = ["--strategy", "wgs", "--method", method, "--" + method, model_path(model), "--genome", fa, "--depth", repr(depth),
            "--seed", str(seed), "--length-min", "100", "--length-max", "20000", "--length-mean", repr(okw["len_mean"]),
            "--length-sd", "400", "--difference-ratio", "%d:%d:%d" % okw["ratio"], "--hp-del-bias", repr(okw["hp_del_bias"]),
            "--accuracy-mean", "%.2f" % okw["accuracy_mean"], "--pass-num", str(okw["pass_num"])]
    ref = R.run_reference(args, logrand=True)
    assert ref["returncode"] == 0, ref["stderr"][-300:]
    o = O.Oracle(method, model_path(model), **okw)
    o.rng_glibc(seed)
    if okw["hp_del_bias"] != 1.0:
        o.hp_bias_prepass([s for _, s in contigs])
    hm = capi.HostModel(H.lib(), capi.host_params(method, **okw), model_path(model))
    blocks = R.split_stats_blocks(ref["stderr"])
    cursor = 0
    for i, (_, s) in enumerate(contigs, start=1):
        o.set_sequence(s, i)
        reads, maf, st = o.simulate_wgs(depth)
        got = ref["files"]["out_%04d.%s" % (i, "bam" if okw["pass_num"] > 1 else "fq.gz")]
        if okw["pass_num"] > 1:
            got = got[got.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n"):]
        assert reads == got, "sequence %d: oracle reads differ from the live reference" % i
        assert maf == ref["files"]["out_%04d.maf.gz" % i]
        assert O.format_stats(st, i) == blocks[i]
        sub = H.run(hm, o.seq_upper(), o.hp(), i, o.bias(), capi.RNG_REPLAY, 0, ref["draws"][cursor:], int(depth * len(s)))
        cursor = o.draws_consumed()
        r2, m2 = H.records_from_events(hm, sub, o.seq_upper(), i)
        assert r2 == got and m2 == maf, "sequence %d: engine core replay differs" % i
    assert cursor == len(ref["draws"])


@pytest.mark.parametrize("k", range(6))
def test_long_transcripts_live(k, tmp_path):
    """transcripts of 20 k .. 600 k bases: start-position tables of high ranks (prob2ssp[rank], rank = ceil(len / 1000),
    :2504-2528, :2855) and table lines of many fgets buffers, against the live reference"""
    rng = np.random.default_rng(1200 + k)
    method = ["qshmm", "errhmm"][k % 2]
    model = "QSHMM-ONT.model" if method == "qshmm" else "ERRHMM-ONT-HQ.model"
    seqset = []
    for t in range(int(rng.integers(2, 6))):
        n = int(rng.integers(20000, 600000))
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes()
        seqset.append(("LONG%d" % t, int(rng.integers(0, 4)), int(rng.integers(0, 4)), s))
    if sum(x[1] + x[2] for x in seqset) == 0:
        seqset[0] = (seqset[0][0], 2, 1, seqset[0][3])
    mean = float(rng.integers(2000, 30000))
    okw = dict(len_min=100, len_max=100000, ratio=(6, 55, 39), hp_del_bias=1.0, len_mean=mean,
               len_sd=float(rng.uniform(0.4, 0.9)) * mean, pass_num=1, accuracy_mean=0.9, accuracy_mean_set=True)
    seed = int(rng.integers(1, 1 << 30))
    inp = str(tmp_path / "input.txt")
    R.write_transcripts(inp, seqset)
    args = ["--strategy", "trans", "--method", method, "--" + method, model_path(model), "--transcript", inp, "--seed",
            str(seed), "--length-mean", repr(okw["len_mean"]), "--length-sd", repr(okw["len_sd"]), "--length-max", "100000",
            "--accuracy-mean", "0.90"]
    ref = R.run_reference(args)
    assert ref["returncode"] == 0, ref["stderr"][-300:]
    o = O.Oracle(method, model_path(model), **okw)
    o.rng_glibc(seed)
    reads, maf, st = o.simulate_set("trans", seqset)
    assert reads == ref["files"]["out.fq.gz"]
    assert maf == ref["files"]["out.maf.gz"]
    assert O.format_stats_set(st) == R.set_stats_block(ref["stderr"])


def _cli_variants():
    qs = ["--strategy", "wgs", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model", "--genome", "tiny.fa"]
    er = ["--strategy", "wgs", "--method", "errhmm", "--errhmm", "ERRHMM-ONT.model", "--genome", "tiny.fa"]
    out = []
    for opt, vals in (("--depth", ["0", "0.0", "-0.5", "abc", "1e3", "1001"]),
                      ("--length-min", ["0", "-1", "1000001", "x"]), ("--length-max", ["0", "1000001", "99"]),
                      ("--length-mean", ["0", "-5", "1000001", "1e7"]), ("--length-sd", ["-1", "1000001"]),
                      ("--accuracy-mean", ["-0.1", "1.01", "abc", "0"]), ("--accuracy-min", ["-1", "2"]),
                      ("--accuracy-max", ["-1", "2"]), ("--difference-ratio", ["1001:1:1", "1:1", "a:b:c", "1:1:1:1", "-1:2:3", "0:0:0"]),
                      ("--pass-num", ["0", "-3", "abc"]), ("--hp-del-bias", ["0", "-1", "11", "abc", "10.5"]),
                      ("--seed", ["abc", "-1"]), ("--id-prefix", ["x" * 200]), ("--prefix", [""])):
        for v in vals:
            out.append(qs + [opt, v])
            if opt in ("--depth", "--pass-num", "--hp-del-bias", "--accuracy-mean"):
                out.append(er + [opt, v])
    out += [["--strategy", "wgsx"] + qs[2:], ["--strategy", "w"] + qs[2:], qs[:2] + ["--method", "qshmmXYZ"] + qs[4:],
            qs[:2] + ["--method", "q"] + qs[4:], ["--strategy", "templ", "--method", "sample", "--template", "tiny.fa"],
            qs + ["--nonsense", "1"], qs + ["--depth"], []]
    return out


@pytest.mark.parametrize("idx", range(len(_cli_variants())))
def test_driver_option_validation_equals_live_reference(idx, tmp_path):
    """every option with out-of-range / malformed values: when the reference rejects the command line, the driver
    prints the same text and exits with the same status (option parsing :257-530, set_sim_param :1451-1688)"""
    import re
    import shutil
    import subprocess
    import __graft_entry__ as G
    args = _cli_variants()[idx]
    shutil.copy(model_path("QSHMM-RSII.model"), tmp_path / "QSHMM-RSII.model")
    shutil.copy(model_path("ERRHMM-ONT.model"), tmp_path / "ERRHMM-ONT.model")
    (tmp_path / "tiny.fa").write_text(">s\nACGTACGTAC\n")
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    pr = subprocess.run([R.REF_BIN] + args, cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert pr.returncode != 0  # (tiny.fa is shorter than 100 bases: nothing gets as far as a simulation)
    G.build_engine()
    pd = subprocess.run([G.build_driver()] + args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)

    def norm(t):
        t = re.sub(r"(?m)^seed : -?\d+$", "seed : <n>", t) if "--seed" not in args else t
        return re.sub(r"(?m)^(\S*pbsim\S*): ", "pbsim: ", t)  # getopt prefixes its messages with argv[0]
    assert pd.returncode == pr.returncode
    assert pd.stdout == pr.stdout
    if not args:  # the help text: the driver's lists its additional engine options
        assert pd.stderr.decode().startswith("\nUSAGE: pbsim [options]") and pr.stderr.decode().startswith("\nUSAGE: pbsim [options]")
        return
    assert norm(pd.stderr.decode()) == norm(pr.stderr.decode())


def test_sample_profile_files_equal_live_reference(tmp_path):
    """--sample + --sample-profile-id stores the filtered reads and their statistics (pbsim.cpp:1317-1326, :590-602);
    --sample-profile-id alone reuses them (:603-615): the driver writes the files the reference writes and prints the
    blocks the reference prints (everything before the first simulation; the driver needs a GPU from there on)"""
    import subprocess
    import __graft_entry__ as G
    rng = np.random.default_rng(17)
    recs = []
    for i in range(60):
        n = int(rng.integers(60, 3000))
        lo = int(rng.integers(2, 15))
        q = bytes(rng.integers(33 + lo, 33 + lo + int(rng.integers(2, 25)), n).astype(np.uint8))
        recs.append(b"@r%d\n" % i + b"G" * n + b"\n+\n" + q + b"\n")
    genome = R.synth_genome(3, [("g", 4000)])[0][1]
    G.build_engine()
    exe = G.build_driver()
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    common = ["--strategy", "wgs", "--method", "sample", "--genome", "genome.fa", "--depth", "2", "--seed", "9",
              "--length-min", "200", "--length-max", "2500", "--accuracy-min", "0.8", "--accuracy-max", "0.97", "--prefix", "out"]
    outs = {}
    for who, binary in (("r", R.REF_BIN), ("d", exe)):
        d = tmp_path / who
        d.mkdir()
        (d / "sample.fq").write_bytes(b"".join(recs))
        R.write_fasta(str(d / "genome.fa"), [("g", genome)])
        p1 = subprocess.run([binary] + common + ["--sample", "sample.fq", "--sample-profile-id", "prof"], cwd=d, env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        p2 = subprocess.run([binary] + common + ["--sample-profile-id", "prof"], cwd=d, env=env, stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, timeout=300)
        p3 = subprocess.run([binary] + common + ["--sample", "sample.fq", "--sample-profile-id", "prof"], cwd=d, env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        cut = ":::: Simulation stats (ref.1) ::::"
        head = lambda t: t.split("ERROR: no usable CUDA device")[0].split(cut)[0]  # noqa: E731
        outs[who] = ((d / "sample_profile_prof.fastq").read_bytes(), (d / "sample_profile_prof.stats").read_bytes(),
                     head(p1.stderr.decode()), head(p2.stderr.decode()), p3.stderr.decode(), p3.returncode)
    assert outs["d"] == outs["r"]
    assert b"\n" in outs["r"][0] and outs["r"][1].startswith(b"num\t")
    assert "exists." in outs["r"][4]


_ODD_FASTA = {
    "empty": b"",
    "header_only": b">s\n",
    "crlf": b">s desc\r\n" + b"ACGT" * 40 + b"\r\n" + b"ACGT" * 40 + b"\r\n",
    "two_headers_empty_first": b">a\n>b\n" + b"ACGT" * 50 + b"\n",
    "short_second": b">a\n" + b"ACGT" * 50 + b"\n>b\nACGT\n",
    "blank_lines": b"\n\n>a\n\n" + b"ACGT" * 50 + b"\n\n",
    "gt_in_seq": b">a\n" + b"ACGT" * 30 + b"\n" + b">" + b"\n" + b"ACGT" * 30 + b"\n",
    "lower_n": b">a\n" + b"acgtnNRY" * 30 + b"\n",
    "long_name": b">" + b"n" * 400 + b"\n" + b"ACGT" * 50 + b"\n",
    "no_header": b"ACGT" * 50 + b"\n",  # the reference crashes on this one; the driver must refuse it
}


@pytest.mark.parametrize("name", sorted(_ODD_FASTA))
def test_driver_handles_odd_fasta_like_the_live_reference(name, tmp_path):
    import subprocess
    import __graft_entry__ as G
    G.build_engine()
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    res = {}
    # "p": the driver's parallel reader (mapped file, a thread per record), which small files normally do not take
    for who, binary in (("r", R.REF_BIN), ("d", G.build_driver()), ("p", G.build_driver())):
        d = tmp_path / who
        d.mkdir()
        (d / "g.fa").write_bytes(_ODD_FASTA[name])
        p = subprocess.run([binary, "--strategy", "wgs", "--method", "qshmm", "--qshmm", model_path("QSHMM-RSII.model"),
                            "--genome", "g.fa", "--depth", "0.3", "--seed", "1", "--prefix", "out"], cwd=d,
                           env=dict(env, PBSIM_INGEST_PARALLEL_MIN="1") if who == "p" else env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        err = p.stderr.decode(errors="replace")
        head = err.split("ERROR: no usable CUDA device")[0].split(":::: Simulation stats (ref.1) ::::")[0]
        refs = {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.endswith(".ref")}
        res[who] = (head, refs, p.returncode)
    assert res["p"] == res["d"]
    if res["r"][2] < 0:  # killed by a signal: undefined behaviour in the reference
        assert res["d"][2] not in (0,) and res["d"][2] > 0
        return
    assert res["d"][0] == res["r"][0]
    assert res["d"][1] == res["r"][1]
    if res["r"][2] != 0:
        assert res["d"][2] == res["r"][2]


_SEQ = b"ACGT" * 100
_ODD_SETS = {
    "ok": (b"T1\t2\t1\t" + _SEQ + b"\nT2\t0\t0\t" + _SEQ + b"\n", "trans"),
    "no_final_newline": (b"T1\t2\t1\t" + _SEQ + b"\nT2\t1\t1\t" + _SEQ, "trans"),
    "empty": (b"", "trans"),
    "non_numeric": (b"T1\tx\ty\t" + _SEQ + b"\n", "trans"),
    "zero_expr_all": (b"T1\t0\t0\t" + _SEQ + b"\n", "trans"),
    "short_seq": (b"T1\t1\t1\tACGTACGTAC\n", "trans"),
    "missing_col": (b"T1\t2\t" + _SEQ + b"\n", "trans"),            # the reference crashes (strtok returns NULL)
    "blank_line": (b"T1\t2\t1\t" + _SEQ + b"\n\nT2\t1\t1\t" + _SEQ + b"\n", "trans"),  # likewise
    "templ_ok": (b">a\n" + _SEQ + b"\n>b\n" + _SEQ + b"\n", "templ"),
    "templ_empty": (b"", "templ"),
    "templ_header_only": (b">a\n", "templ"),
    "templ_blank": (b">a\n\n" + _SEQ + b"\n\n>b\n" + _SEQ + b"\n", "templ"),
}


@pytest.mark.parametrize("name", sorted(_ODD_SETS))
def test_driver_handles_odd_tables_like_the_live_reference(name, tmp_path):
    import subprocess
    import __graft_entry__ as G
    G.build_engine()
    data, strategy = _ODD_SETS[name]
    env = dict(os.environ, PATH=R.SHIMS + ":" + os.environ.get("PATH", ""))
    res = {}
    for who, binary in (("r", R.REF_BIN), ("d", G.build_driver())):
        d = tmp_path / who
        d.mkdir()
        (d / "in.txt").write_bytes(data)
        p = subprocess.run([binary, "--strategy", strategy, "--method", "qshmm", "--qshmm", model_path("QSHMM-RSII.model"),
                            "--transcript" if strategy == "trans" else "--template", "in.txt", "--seed", "1", "--prefix", "out",
                            "--length-mean", "200", "--length-sd", "100"], cwd=d, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=120)
        err = p.stderr.decode(errors="replace")
        res[who] = (err.split("ERROR: no usable CUDA device")[0].split(":::: Simulation stats ::::")[0], p.returncode)
    if res["r"][1] < 0:  # the reference died of a signal: the driver must stop with a message instead
        assert res["d"][1] > 0 and "ERROR:" in res["d"][0]
        return
    assert res["d"][0] == res["r"][0]
