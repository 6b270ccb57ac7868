"""Helpers shared by the parity tests: load tests/golden/<case> (written by oracle/make_golden.py
from the unmodified reference binary) and drive the CPU oracle over the same inputs."""
import gzip
import json
import os
import tempfile

import numpy as np

from oracle import oracle as O
from oracle import refrun as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)
_MODEL_CACHE = os.path.join(tempfile.gettempdir(), "pbsim_b200_models_%d" % os.getuid())


def model_path(name):
    """The reference's data/*.model files are input data; tests carry them gz-compressed under
    tests/golden/models/ (packed by oracle/make_golden.py) and unpack them on first use."""
    os.makedirs(_MODEL_CACHE, exist_ok=True)
    p = os.path.join(_MODEL_CACHE, name)
    src = os.path.join(GOLDEN, "models", name + ".gz")
    if not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src):
        with gzip.open(src, "rb") as f, open(p + ".tmp%d" % os.getpid(), "wb") as g:
            g.write(f.read())
        os.replace(p + ".tmp%d" % os.getpid(), p)
    return p


def case_names():
    return sorted(d for d in os.listdir(GOLDEN)
                  if os.path.isdir(os.path.join(GOLDEN, d)) and d not in ("models", "stats", "sets", "sample"))


def gz_read(path):
    with gzip.open(path, "rb") as f:
        return f.read()


class Case:
    def __init__(self, name):
        self.name = name
        self.dir = os.path.join(GOLDEN, name)
        with open(os.path.join(self.dir, "case.json")) as f:
            self.meta = json.load(f)
        self.method = self.meta["method"]
        self.model = model_path(self.meta["model"])
        self.depth = self.meta["depth"]
        self.seed = self.meta["seed"]
        self.pass_num = self.meta["pass_num"]
        self.okw = dict(self.meta["oracle_kwargs"])
        if "ratio" in self.okw:
            self.okw["ratio"] = tuple(self.okw["ratio"])
        self.contigs = R.read_fasta(os.path.join(self.dir, "genome.fa.gz"))
        with open(os.path.join(self.dir, "stderr.txt")) as f:
            self.stderr = f.read()
        self.stats_blocks = R.split_stats_blocks(self.stderr)
        self.marks = np.load(os.path.join(self.dir, "marks.npy"))
        with open(os.path.join(self.dir, "ndraws.txt")) as f:
            self.ndraws = int(f.read())

    def reads(self, i):
        """Reference FASTQ text, or SAM records WITHOUT the two header lines main() writes (pbsim.cpp:721-722)."""
        data = gz_read(os.path.join(self.dir, "seq%d.reads.gz" % i))
        if self.pass_num > 1:
            end = data.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n")
            data = data[end:]
        return data

    def sam_header(self, i):
        data = gz_read(os.path.join(self.dir, "seq%d.reads.gz" % i))
        end = data.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n")
        return data[:end]

    def maf(self, i):
        return gz_read(os.path.join(self.dir, "seq%d.maf.gz" % i))

    def new_oracle(self):
        return O.Oracle(self.method, self.model, **self.okw)

    def run_oracle(self, rng="glibc", log=None):
        """Runs every sequence; returns (list of per-sequence dicts, oracle)."""
        o = self.new_oracle()
        if rng == "glibc":
            o.rng_glibc(self.seed)
        elif rng == "replay":
            o.rng_replay(log)
        else:
            o.rng_philox(self.seed)
        if self.okw.get("hp_del_bias", 1.0) != 1.0:
            o.hp_bias_prepass([s for _, s in self.contigs])
        out = []
        for i, (_, s) in enumerate(self.contigs, start=1):
            o.set_sequence(s, i)
            reads, maf, st = o.simulate_wgs(self.depth)
            out.append(dict(reads=reads, maf=maf, stats=st, stats_text=O.format_stats(st, i),
                            info=o.readinfo(), bias=o.bias(), freq_len=o.freq_len(),
                            freq_accuracy=o.freq_accuracy(), draws_end=o.draws_consumed()))
        return out, o


def set_case_names():
    return sorted(os.listdir(os.path.join(GOLDEN, "sets")))


def read_transcripts(path):
    """The 4-column table as get_transcript_inf / simulate_by_*_trans see it (pbsim.cpp:1095-1120, :2748-2772):
    id (truncated to 128 chars), plus, minus, sequence."""
    out = []
    with gzip.open(path, "rb") as f:
        for line in f:
            name, plus, minus, seq = line.rstrip(b"\n").split(b"\t")
            out.append((name[:128].decode(), int(plus), int(minus), seq))
    return out


class SetCase:
    """tests/golden/sets/<name>: a transcript (--strategy trans) or template (--strategy templ) run of the reference"""

    def __init__(self, name):
        self.name = name
        self.dir = os.path.join(GOLDEN, "sets", name)
        with open(os.path.join(self.dir, "case.json")) as f:
            self.meta = json.load(f)
        self.strategy = self.meta["strategy"]
        self.method = self.meta["method"]
        self.model = model_path(self.meta["model"])
        self.seed = self.meta["seed"]
        self.pass_num = self.meta["pass_num"]
        self.okw = dict(self.meta["oracle_kwargs"])
        if "ratio" in self.okw:
            self.okw["ratio"] = tuple(self.okw["ratio"])
        inp = os.path.join(self.dir, "input.txt.gz")
        if self.strategy == "trans":
            self.seqset = read_transcripts(inp)
        else:
            self.seqset = [(n, 1, 0, s) for n, s in R.read_fasta(inp)]
        with open(os.path.join(self.dir, "stderr.txt")) as f:
            self.stderr = f.read()
        self.stats_text = R.set_stats_block(self.stderr)
        self.marks = np.load(os.path.join(self.dir, "marks.npy"))
        with open(os.path.join(self.dir, "ndraws.txt")) as f:
            self.ndraws = int(f.read())

    def reads(self):
        data = gz_read(os.path.join(self.dir, "reads.gz"))
        if self.pass_num > 1:
            end = data.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n")
            data = data[end:]
        return data

    def sam_header(self):
        data = gz_read(os.path.join(self.dir, "reads.gz"))
        return data[:data.index(b"PM:SEQUELII\n") + len(b"PM:SEQUELII\n")]

    def maf(self):
        return gz_read(os.path.join(self.dir, "maf.gz"))

    def run_oracle(self, rng="glibc", log=None):
        o = O.Oracle(self.method, self.model, **self.okw)
        if rng == "glibc":
            o.rng_glibc(self.seed)
        elif rng == "replay":
            o.rng_replay(log)
        else:
            o.rng_philox(self.seed)
        reads, maf, st = o.simulate_set(self.strategy, self.seqset)
        return dict(reads=reads, maf=maf, stats=st, stats_text=O.format_stats_set(st), info=o.readinfo(),
                    bias=o.bias(), freq_len=o.freq_len(), freq_accuracy=o.freq_accuracy(),
                    draws_end=o.draws_consumed()), o


def sample_case_names():
    return sorted(os.listdir(os.path.join(GOLDEN, "sample")))


class SampleCase:
    """tests/golden/sample/<name>: a --method sample run of the reference; the sample FASTQ is the reference's own
    output of another golden case (meta["sample_of"])"""

    def __init__(self, name):
        self.name = name
        self.dir = os.path.join(GOLDEN, "sample", name)
        with open(os.path.join(self.dir, "case.json")) as f:
            self.meta = json.load(f)
        self.depth = self.meta["depth"]
        self.seed = self.meta["seed"]
        self.okw = dict(self.meta["oracle_kwargs"])
        if "ratio" in self.okw:
            self.okw["ratio"] = tuple(self.okw["ratio"])
        self.contigs = R.read_fasta(os.path.join(self.dir, "genome.fa.gz"))
        src = os.path.join(GOLDEN, self.meta["sample_of"])
        self.sample_fastq = b"".join(gz_read(os.path.join(src, "seq%d.reads.gz" % i))
                                     for i in range(1, 10) if os.path.exists(os.path.join(src, "seq%d.reads.gz" % i)))
        self.pool = O.sample_pool(self.sample_fastq, **self.meta["pool_kwargs"])
        with open(os.path.join(self.dir, "stderr.txt")) as f:
            self.stderr = f.read()
        self.stats_blocks = R.split_stats_blocks(self.stderr)
        self.marks = np.load(os.path.join(self.dir, "marks.npy"))
        with open(os.path.join(self.dir, "ndraws.txt")) as f:
            self.ndraws = int(f.read())

    def reads(self, i):
        return gz_read(os.path.join(self.dir, "seq%d.reads.gz" % i))

    def maf(self, i):
        return gz_read(os.path.join(self.dir, "seq%d.maf.gz" % i))
