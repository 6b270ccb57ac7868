#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate tests/golden/ from the UNMODIFIED reference binary.

Run in the build container (where /root/reference exists):
    make -C oracle ref && python oracle/make_golden.py

For every case below it writes tests/golden/<case>/
    case.json        parameters, the exact reference command line, toolchain stamp
    genome.fa.gz     the synthetic genome given to the reference
    seqN.reads.gz    what the reference wrote to <prefix>_NNNN.fq.gz (FASTQ) or .bam (SAM text, header included)
    seqN.maf.gz      what it wrote to <prefix>_NNNN.maf.gz
    seqN.ref.gz      <prefix>_NNNN.ref
    stderr.txt       the reference's stderr (parameter / reference / simulation stats blocks)
    marks.npy        draw index after each (read, pass), from the interposed rand() log
    ndraws.txt       total rand() calls
    rand_head.npy    first 64 draws (pins the glibc rand() restatement for this seed)
The draw log itself is not committed (it is regenerated from the seed by oracle/glibc_rand.c).
"""
import gzip
import hashlib
import json
import os
import platform
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refrun as R  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
DATA = R.REF_DATA

G_PLAIN = dict(seed=1, contigs=[("chrA", 20000), ("chrB", 5000)])
G_QUIRK = dict(seed=2, contigs=[("c1", 30000), ("c2", 150), ("c3", 8000)], n_runs=6, hp_plants=40,
               lowercase_frac=0.1, iupac=5)

G_LONGHP = dict(seed=3, contigs=[("h1", 20000)], long_runs=[12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 34, 38, 40, 46] * 3)

G_HP11 = dict(seed=4, contigs=[("h11", 30000)], long_runs=[11, 13, 15, 17, 11, 13, 21, 11, 12, 14] * 12, n_runs=8, iupac=12)

CASES = {
    # name: (method, model, genome spec, depth, seed, extra CLI args, oracle kwargs)
    "qs_rsii_basic": ("qshmm", "QSHMM-RSII.model", G_PLAIN, 5, 42,
                      ["--length-mean", "1000", "--length-sd", "700"],
                      dict(len_mean=1000.0, len_sd=700.0)),
    "qs_rsii_quirks": ("qshmm", "QSHMM-RSII.model", G_QUIRK, 4, 7,
                       ["--length-mean", "800", "--length-sd", "600"],
                       dict(len_mean=800.0, len_sd=600.0)),
    "qs_ont_hpbias": ("qshmm", "QSHMM-ONT.model", G_QUIRK, 3, 9,
                      ["--length-mean", "800", "--length-sd", "600", "--hp-del-bias", "4",
                       "--difference-ratio", "39:24:36"],
                      dict(len_mean=800.0, len_sd=600.0, hp_del_bias=4.0, ratio=(39, 24, 36))),
    "err_onthq_basic": ("errhmm", "ERRHMM-ONT-HQ.model", G_QUIRK, 4, 11,
                        ["--length-mean", "800", "--length-sd", "600"],
                        dict(len_mean=800.0, len_sd=600.0)),
    "err_sequel_hiacc": ("errhmm", "ERRHMM-SEQUEL.model", G_QUIRK, 4, 13,
                         ["--length-mean", "800", "--length-sd", "600", "--accuracy-mean", "0.98",
                          "--hp-del-bias", "2.5"],
                         dict(len_mean=800.0, len_sd=600.0, accuracy_mean=0.98, accuracy_mean_set=True,
                              hp_del_bias=2.5)),
    "err_sequel_multipass": ("errhmm", "ERRHMM-SEQUEL.model", G_PLAIN, 2, 5,
                             ["--length-mean", "800", "--length-sd", "600", "--pass-num", "3"],
                             dict(len_mean=800.0, len_sd=600.0, pass_num=3)),
    "qs_rsii_multipass": ("qshmm", "QSHMM-RSII.model", G_PLAIN, 2, 6,
                          ["--length-mean", "800", "--length-sd", "600", "--pass-num", "2", "--id-prefix", "XY"],
                          dict(len_mean=800.0, len_sd=600.0, pass_num=2, id_prefix="XY")),
    # every base QV 0, deletions dominate: runs of >= 15 deletions after one read position (continuation entries)
    "qs_delheavy_uniform": ("qshmm", "QSHMM-RSII.model", G_PLAIN, 8, 31,
                            ["--length-mean", "1500", "--length-sd", "900", "--accuracy-mean", "0.0",
                             "--difference-ratio", "1:1:1000"],
                            dict(len_mean=1500.0, len_sd=900.0, accuracy_mean=0.0, accuracy_mean_set=True,
                                 ratio=(1, 1, 1000))),
    # the same through the homopolymer-bias path: long even-length homopolymers are deleted wholesale
    "qs_delheavy_bias": ("qshmm", "QSHMM-RSII.model", G_LONGHP, 6, 32,
                         ["--length-mean", "1500", "--length-sd", "900", "--accuracy-mean", "0.3",
                          "--difference-ratio", "1:1:1000", "--hp-del-bias", "10"],
                         dict(len_mean=1500.0, len_sd=900.0, accuracy_mean=0.3, accuracy_mean_set=True,
                              ratio=(1, 1, 1000), hp_del_bias=10.0)),
    # default bias, many homopolymers >= 11 and N runs, deletion-rich: exercises the deletion-run repair of the
    # segment-parallel path (a deletion cannot follow a base of an 11/13/15..-run: hp_del_bias[11] is 0)
    "qs_hp11_uniform": ("qshmm", "QSHMM-RSII.model", G_HP11, 6, 33,
                        ["--length-mean", "3000", "--length-sd", "1500", "--accuracy-mean", "0.7",
                         "--difference-ratio", "10:20:70"],
                        dict(len_mean=3000.0, len_sd=1500.0, accuracy_mean=0.7, accuracy_mean_set=True,
                             ratio=(10, 20, 70))),
    "qs_rsii_fixedlen": ("qshmm", "QSHMM-RSII.model", G_PLAIN, 3, 21,
                         ["--length-mean", "500", "--length-sd", "0", "--accuracy-mean", "0.9",
                          "--length-min", "50", "--length-max", "5000"],
                         dict(len_mean=500.0, len_sd=0.0, accuracy_mean=0.9, accuracy_mean_set=True,
                              len_min=50, len_max=5000)),
}


# transcript / template strategies: name: (strategy, method, model, synth_set kwargs, seed, extra CLI args, oracle kwargs)
SET_CASES = {
    "tr_qs_rsii_basic": ("trans", "qshmm", "QSHMM-RSII.model", dict(seed=3, n=12), 3, [], {}),
    "tr_err_onthq_basic": ("trans", "errhmm", "ERRHMM-ONT-HQ.model", dict(seed=4, n=12), 4, [], {}),
    "tm_qs_rsii_basic": ("templ", "qshmm", "QSHMM-RSII.model", dict(seed=5, n=12, long_every=4), 5, [], {}),
    # lower-case sequences: only simulate_by_qshmm_trans upper-cases the first base (:2774 vs :3330, :4474)
    "tm_err_sequel_multipass": ("templ", "errhmm", "ERRHMM-SEQUEL.model",
                                dict(seed=10, n=10, lowercase_first=0.5, hp_plants=2), 10,
                                ["--pass-num", "3"], dict(pass_num=3)),
    "tr_qs_ont_hpbias": ("trans", "qshmm", "QSHMM-ONT.model",
                         dict(seed=7, n=12, hp_plants=3, lowercase_first=0.5, iupac=0.01), 7,
                         ["--hp-del-bias", "3", "--difference-ratio", "39:24:36"],
                         dict(hp_del_bias=3.0, ratio=(39, 24, 36))),
    "tr_err_ont_hpbias": ("trans", "errhmm", "ERRHMM-ONT.model",
                          dict(seed=8, n=12, hp_plants=3, lowercase_first=0.5, iupac=0.01), 8,
                          ["--hp-del-bias", "2"], dict(hp_del_bias=2.0)),
    "tm_qs_rsii_quirks": ("templ", "qshmm", "QSHMM-RSII.model",
                          dict(seed=9, n=12, lowercase_first=0.5, hp_plants=3, iupac=0.01), 9, [], {}),
    # transcripts longer than one fgets buffer (BUF_SIZE 10240), two passes, short length distribution
    "tr_qs_rsii_multipass_long": ("trans", "qshmm", "QSHMM-RSII.model", dict(seed=11, n=8, long_every=4), 11,
                                  ["--pass-num", "2", "--length-mean", "2500", "--length-sd", "2000",
                                   "--id-prefix", "Q"],
                                  dict(pass_num=2, len_mean=2500.0, len_sd=2000.0, id_prefix="Q")),
    "tm_qs_rsii_hpbias": ("templ", "qshmm", "QSHMM-RSII.model", dict(seed=12, n=12, hp_plants=3), 12,
                          ["--hp-del-bias", "3"], dict(hp_del_bias=3.0)),
    # accuracy-100 reads in simulate_by_errhmm_trans: the copy loop shares its variable with the transcript's read
    # counter (:4487, :4532), so the transcript's remaining reads are (usually) never simulated
    "tr_err_sequel_acc100": ("trans", "errhmm", "ERRHMM-SEQUEL.model", dict(seed=13, n=14, max_exp=9), 13,
                             ["--accuracy-mean", "0.98"], dict(accuracy_mean=0.98, accuracy_mean_set=True)),
}


# --method sample (pbsim.cpp:1694): the sample FASTQ is the reference's own output of the case named in `sample_of`
SAMPLE_CASES = {
    "sample_basic": dict(sample_of="qs_rsii_basic", genome=dict(seed=5, contigs=[("g1", 30000), ("g2", 900)], n_runs=3,
                                                              hp_plants=20),
                         depth=5, seed=3, extra=[], okw={}, pool={}),
    # several copies per pool entry (the quality buffer is cut at the end of every read), deletion-rich, bias, filters
    "sample_quirks": dict(sample_of="qs_rsii_basic", genome=dict(seed=6, contigs=[("g1", 25000), ("g2", 700)], n_runs=4,
                                                               hp_plants=30, iupac=6),
                          depth=14, seed=5,
                          extra=["--hp-del-bias", "3", "--difference-ratio", "20:30:50", "--length-min", "300",
                                 "--length-max", "2500", "--accuracy-min", "0.78", "--accuracy-max", "0.93"],
                          okw=dict(hp_del_bias=3.0, ratio=(20, 30, 50), len_min=300, len_max=2500),
                          pool=dict(len_min=300, len_max=2500, accuracy_min=0.78, accuracy_max=0.93)),
}


STATS_CASES = {
    "qs_rsii_len3k": ("qshmm", "QSHMM-RSII.model", ["--length-mean", "3000", "--length-sd", "2300"]),
    "err_onthq_len3k": ("errhmm", "ERRHMM-ONT-HQ.model", ["--length-mean", "3000", "--length-sd", "2300"]),
}


def toolchain_stamp():
    gxx = subprocess.run(["g++", "--version"], stdout=subprocess.PIPE).stdout.decode().splitlines()[0]
    libc = " ".join(platform.libc_ver())
    with open(R.REF_SRC, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    return {"gxx": gxx, "libc": libc, "flags": "-O2", "reference_sha256": sha, "machine": platform.machine()}


def gz_write(path, data):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(data)


def main():
    assert R.build_reference(), "reference source not present: run this in the build container"
    stamp = toolchain_stamp()
    # `--set NAME`: (re)generate one transcript / template case only (the others stay as committed)
    only_set = sys.argv[sys.argv.index("--set") + 1] if "--set" in sys.argv else None
    for name, (method, model, gspec, depth, seed, extra, okw) in ([] if only_set else CASES.items()):
        d = os.path.join(GOLDEN, name)
        os.makedirs(d, exist_ok=True)
        contigs = R.synth_genome(**gspec)
        fa = os.path.join(d, "genome.fa")
        R.write_fasta(fa, contigs)
        args = ["--strategy", "wgs", "--method", method, "--" + method, os.path.join(DATA, model),
                "--genome", fa, "--depth", str(depth), "--seed", str(seed)] + extra
        plain = R.run_reference(args, logrand=False)
        logged = R.run_reference(args, logrand=True)
        assert plain["returncode"] == 0, plain["stderr"]
        assert plain["files"] == logged["files"], "interposed rand() changed the reference's output"
        pass_num = okw.get("pass_num", 1)
        for i in range(1, len(contigs) + 1):
            reads = plain["files"]["out_%04d.%s" % (i, "fq.gz" if pass_num == 1 else "bam")]
            gz_write(os.path.join(d, "seq%d.reads.gz" % i), reads)
            gz_write(os.path.join(d, "seq%d.maf.gz" % i), plain["files"]["out_%04d.maf.gz" % i])
            gz_write(os.path.join(d, "seq%d.ref.gz" % i), plain["files"]["out_%04d.ref" % i])
        with open(fa, "rb") as f:
            gz_write(fa + ".gz", f.read())
        os.remove(fa)
        stderr = plain["stderr"].replace(fa, "genome.fa")
        with open(os.path.join(d, "stderr.txt"), "w") as f:
            f.write(stderr)
        np.save(os.path.join(d, "marks.npy"), logged["marks"])
        np.save(os.path.join(d, "rand_head.npy"), logged["draws"][:64])
        with open(os.path.join(d, "ndraws.txt"), "w") as f:
            f.write("%d\n" % len(logged["draws"]))
        case = dict(name=name, method=method, model=model, depth=depth, seed=seed, extra_args=extra,
                    oracle_kwargs=okw, genome_spec=gspec, n_seq=len(contigs), pass_num=pass_num,
                    toolchain=stamp,
                    command="pbsim --strategy wgs --method %s --%s data/%s --genome genome.fa --depth %s --seed %d %s"
                            % (method, method, model, depth, seed, " ".join(extra)))
        with open(os.path.join(d, "case.json"), "w") as f:
            json.dump(case, f, indent=1, sort_keys=True)
        print("golden", name, {k: len(v) for k, v in plain["files"].items() if not k.endswith(".ref")},
              "draws", len(logged["draws"]))

    for name, (strategy, method, model, skw, seed, extra, okw) in SET_CASES.items():
        if only_set and name != only_set:
            continue
        d = os.path.join(GOLDEN, "sets", name)
        os.makedirs(d, exist_ok=True)
        seqset = R.synth_set(**skw)
        fn = os.path.join(d, "input.txt")
        (R.write_transcripts if strategy == "trans" else R.write_templates)(fn, seqset)
        args = ["--strategy", strategy, "--method", method, "--" + method, os.path.join(DATA, model),
                "--transcript" if strategy == "trans" else "--template", fn, "--seed", str(seed)] + extra
        plain = R.run_reference(args, logrand=False)
        logged = R.run_reference(args, logrand=True)
        assert plain["returncode"] == 0, plain["stderr"]
        assert plain["files"] == logged["files"], "interposed rand() changed the reference's output"
        pass_num = okw.get("pass_num", 1)
        gz_write(os.path.join(d, "reads.gz"), plain["files"]["out.fq.gz" if pass_num == 1 else "out.bam"])
        gz_write(os.path.join(d, "maf.gz"), plain["files"]["out.maf.gz"])
        with open(fn, "rb") as f:
            gz_write(fn + ".gz", f.read())
        os.remove(fn)
        with open(os.path.join(d, "stderr.txt"), "w") as f:
            f.write(plain["stderr"].replace(fn, "input.txt"))
        np.save(os.path.join(d, "marks.npy"), logged["marks"])
        with open(os.path.join(d, "ndraws.txt"), "w") as f:
            f.write("%d\n" % len(logged["draws"]))
        case = dict(name=name, strategy=strategy, method=method, model=model, seed=seed, extra_args=extra,
                    oracle_kwargs=okw, set_spec=skw, n_seq=len(seqset), pass_num=pass_num, toolchain=stamp,
                    command="pbsim --strategy %s --method %s --%s data/%s --%s input.txt --seed %d %s"
                            % (strategy, method, method, model, "transcript" if strategy == "trans" else "template",
                               seed, " ".join(extra)))
        with open(os.path.join(d, "case.json"), "w") as f:
            json.dump(case, f, indent=1, sort_keys=True)
        print("golden set", name, {k: len(v) for k, v in plain["files"].items()}, "draws", len(logged["draws"]))

    if only_set:
        return
    for name, sc in SAMPLE_CASES.items():
        d = os.path.join(GOLDEN, "sample", name)
        os.makedirs(d, exist_ok=True)
        src = os.path.join(GOLDEN, sc["sample_of"])
        fq = b""
        for i in range(1, 10):
            pth = os.path.join(src, "seq%d.reads.gz" % i)
            if os.path.exists(pth):
                with gzip.open(pth, "rb") as f:
                    fq += f.read()
        sfq = os.path.join(d, "sample.fq")
        with open(sfq, "wb") as f:
            f.write(fq)
        contigs = R.synth_genome(**sc["genome"])
        fa = os.path.join(d, "genome.fa")
        R.write_fasta(fa, contigs)
        args = ["--strategy", "wgs", "--method", "sample", "--sample", sfq, "--genome", fa, "--depth", str(sc["depth"]),
                "--seed", str(sc["seed"])] + sc["extra"]
        plain = R.run_reference(args, logrand=False)
        logged = R.run_reference(args, logrand=True)
        assert plain["returncode"] == 0, plain["stderr"]
        assert plain["files"] == logged["files"]
        for i in range(1, len(contigs) + 1):
            gz_write(os.path.join(d, "seq%d.reads.gz" % i), plain["files"]["out_%04d.fq.gz" % i])
            gz_write(os.path.join(d, "seq%d.maf.gz" % i), plain["files"]["out_%04d.maf.gz" % i])
        with open(fa, "rb") as f:
            gz_write(fa + ".gz", f.read())
        os.remove(fa)
        os.remove(sfq)
        with open(os.path.join(d, "stderr.txt"), "w") as f:
            f.write(plain["stderr"].replace(fa, "genome.fa").replace(sfq, "sample.fq"))
        np.save(os.path.join(d, "marks.npy"), logged["marks"])
        with open(os.path.join(d, "ndraws.txt"), "w") as f:
            f.write("%d\n" % len(logged["draws"]))
        case = dict(name=name, method="sample", sample_of=sc["sample_of"], depth=sc["depth"], seed=sc["seed"],
                    extra_args=sc["extra"], oracle_kwargs=sc["okw"], pool_kwargs=sc["pool"], genome_spec=sc["genome"],
                    n_seq=len(contigs), toolchain=stamp,
                    command="pbsim --strategy wgs --method sample --sample sample.fq --genome genome.fa --depth %s --seed %d %s"
                            % (sc["depth"], sc["seed"], " ".join(sc["extra"])))
        with open(os.path.join(d, "case.json"), "w") as f:
            json.dump(case, f, indent=1, sort_keys=True)
        print("golden sample", name, {k: len(v) for k, v in plain["files"].items() if not k.endswith(".ref")},
              "draws", len(logged["draws"]))

    # command-line validation: what the reference prints and returns for invocations it rejects before simulating
    # (set_sim_param :1451-1688, get_genome_inf :896-991, the file openers); the driver must say the same
    import shutil
    import tempfile
    work = tempfile.mkdtemp(prefix="pbsim_cli_")
    shutil.copy(os.path.join(DATA, "QSHMM-RSII.model"), os.path.join(work, "QSHMM-RSII.model"))
    shutil.copy(os.path.join(DATA, "ERRHMM-ONT.model"), os.path.join(work, "ERRHMM-ONT.model"))
    with open(os.path.join(work, "tiny.fa"), "w") as f:
        f.write(">s\nACGTACGTAC\n")
    with open(os.path.join(work, "tiny.fq"), "w") as f:
        f.write("@r1\nACGT\n+\nIIII\n@r2\nACGTA\n+\nIIIII\n")
    with open(os.path.join(work, "nonl.fq"), "w") as f:      # last record without line feed: not counted (:1218)
        f.write("@r1\nACGTAC\n+\nIIIIII\n@r2\nACGTA\n+\n55555")
    with open(os.path.join(work, "long.fq"), "w") as f:      # lines longer than one fgets buffer (BUF_SIZE 10240, :20)
        f.write("@r1\n" + "ACGT" * 6250 + "\n+\n" + "5I+?" * 6250 + "\n@r2\n" + "A" * 10239 + "\n+\n" + "9" * 10239 + "\n"
                + "@r3\nACGT\n+\n!!!!\n")
    sm = ["--strategy", "wgs", "--method", "sample", "--genome", "tiny.fa"]
    qs = ["--strategy", "wgs", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model", "--genome", "tiny.fa"]
    cli_cases = [
        ["--strategy", "foo"], ["--strategy", "wgs"], ["--strategy", "wgs", "--method", "qshmm"],
        ["--strategy", "wgs", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model"],
        ["--strategy", "wgs", "--method", "bar", "--genome", "tiny.fa"],
        ["--strategy", "wgs", "--method", "errhmm", "--genome", "tiny.fa"],
        qs + ["--depth", "-1"], qs + ["--length-min", "500", "--length-max", "100"], qs + ["--difference-ratio", "1:2"],
        qs + ["--difference-ratio", "0:0:0"], qs + ["--accuracy-mean", "1.5"], qs + ["--pass-num", "0"],
        qs + ["--hp-del-bias", "0"], qs + ["--length-mean", "0"], qs + ["--length-sd", "-3"], qs + ["--length-max", "2000000"],
        qs + ["--seed", "-5"], qs,
        ["--strategy", "wgs", "--method", "qshmm", "--qshmm", "nofile.model", "--genome", "tiny.fa"],
        ["--strategy", "wgs", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model", "--genome", "nofile.fa"],
        ["--strategy", "trans", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model"],
        ["--strategy", "trans", "--method", "qshmm", "--qshmm", "QSHMM-RSII.model", "--transcript", "nofile.tsv"],
        ["--strategy", "templ", "--method", "errhmm", "--errhmm", "ERRHMM-ONT.model"],
        ["--strategy", "templ", "--method", "errhmm", "--errhmm", "ERRHMM-ONT.model", "--template", "nofile.fa"],
        ["--strategy", "trans", "--method", "sample", "--transcript", "nofile.tsv"],
        sm, sm + ["--sample", "nofile.fq"], sm + ["--sample", "tiny.fq", "--pass-num", "2"],
        sm + ["--sample-profile-id", "nosuchprofile"], sm + ["--sample", "tiny.fq"],
        sm + ["--sample", "tiny.fq", "--accuracy-min", "1.5"],
        sm + ["--sample", "tiny.fq", "--length-min", "1"], sm + ["--sample", "nonl.fq", "--length-min", "1"],
        sm + ["--sample", "long.fq", "--length-min", "1"], sm + ["--sample", "long.fq", "--length-min", "1", "--accuracy-min", "0.5"],
        sm + ["--sample", "long.fq", "--length-min", "1", "--accuracy-max", "0.4"],
        sm + ["--sample", "long.fq", "--length-min", "1", "--length-max", "10239", "--accuracy-min", "0"],
    ]
    cli_out = []
    for a in cli_cases:
        p = subprocess.run([R.REF_BIN] + a, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        cli_out.append(dict(args=a, returncode=p.returncode, stderr=p.stderr.decode(), stdout=p.stdout.decode()))
    with open(os.path.join(GOLDEN, "cli_errors.json"), "w") as f:
        json.dump(dict(toolchain=stamp, cases=cli_out), f, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    print("cli error cases", len(cli_out))

    # distribution fixtures for the PHILOX-mode statistical parity tests: larger reference runs, reduced to
    # histograms (tests/stats_util.py) so that only a few KB are committed
    from tests import stats_util as SU
    sdir = os.path.join(GOLDEN, "stats")
    os.makedirs(sdir, exist_ok=True)
    import tempfile
    for sname, (method, model, extra) in STATS_CASES.items():
        tmp = tempfile.mkdtemp()
        fa = os.path.join(tmp, "g.fa")
        R.write_fasta(fa, R.synth_genome(77, [("s1", 1500000)]))
        args = ["--strategy", "wgs", "--method", method, "--" + method, os.path.join(DATA, model), "--genome", fa,
                "--depth", "8", "--seed", "2024"] + extra
        res = R.run_reference(args)
        assert res["returncode"] == 0, res["stderr"]
        st = SU.parse_outputs(res["files"]["out_0001.fq.gz"], res["files"]["out_0001.maf.gz"])
        np.savez_compressed(os.path.join(sdir, sname + ".npz"), lengths=st["lengths"].astype(np.int32),
                            accuracy=st["accuracy"].astype(np.float32), err_accuracy=st["err_accuracy"].astype(np.float32),
                            qv_hist=st["qv_hist"], events=st["events"], per_read=st["per_read"].astype(np.int32),
                            plus=st["plus"], n=st["n"])
        with open(os.path.join(sdir, sname + ".json"), "w") as f:
            json.dump(dict(method=method, model=model, extra_args=extra, depth=8, seed=2024, genome_bp=1500000,
                           toolchain=stamp), f, indent=1)
        print("stats", sname, st["n"], "reads", st["events"].tolist())

    # the reference's data/*.model files are INPUT DATA of the path (HMM parameters), not code;
    # carried gz-compressed so that the -m gpu tests and bench.py can run where /root/reference is absent
    mdir = os.path.join(GOLDEN, "models")
    os.makedirs(mdir, exist_ok=True)
    for fn in sorted(os.listdir(DATA)):
        if fn.endswith(".model"):
            with open(os.path.join(DATA, fn), "rb") as f:
                gz_write(os.path.join(mdir, fn + ".gz"), f.read())

    # glibc rand() known answers taken from the system libc here
    import ctypes
    libc = ctypes.CDLL(None)
    kat = {}
    for seed in (0, 1, 42, 2024, 4294967295):
        libc.srand(ctypes.c_uint(seed))
        kat[str(seed)] = [int(libc.rand()) for _ in range(400)]
    with open(os.path.join(GOLDEN, "glibc_rand_kat.json"), "w") as f:
        json.dump({"libc": stamp["libc"], "values": kat}, f)


if __name__ == "__main__":
    main()
