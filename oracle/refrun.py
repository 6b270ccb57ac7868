"""TEST INFRASTRUCTURE — run the UNMODIFIED reference binary (oracle/_ref/pbsim*), built by
`make -C oracle ref` from /root/reference/src/pbsim.cpp, and collect what it wrote.

The reference pipes its text through popen("gzip > f") / popen("samtools view -b -o f -")
(pbsim.cpp:708-730); PATH shims (oracle/_ref/shims) turn both into `cat`, so the files
hold the uncompressed FASTQ / SAM / MAF text the generator produced.
"""
import glob
import os
import shutil
import subprocess
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "pbsim")
REF_BIN_LOG = os.path.join(REF_DIR, "pbsim_logrand")
SHIMS = os.path.join(REF_DIR, "shims")
REF_SRC = "/root/reference/src/pbsim.cpp"
REF_DATA = "/root/reference/data"


def have_reference_binary():
    return os.path.exists(REF_BIN) and os.path.exists(os.path.join(SHIMS, "gzip"))


def build_reference():
    """Compile the reference where it lies (only possible where /root/reference exists)."""
    if not os.path.exists(REF_SRC):
        return False
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return True


def synth_genome(seed, contigs, n_runs=0, hp_plants=0, lowercase_frac=0.0, iupac=0, long_runs=()):
    """Seeded synthetic genome: i.i.d. ACGT, optional N runs, planted homopolymers (2..15),
    lower-case stretches and isolated IUPAC codes.  Returns [(name, bytes)]."""
    rng = np.random.default_rng(seed)
    out = []
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    for name, n in contigs:
        s = alphabet[rng.integers(0, 4, size=n)].copy()
        for _ in range(hp_plants):
            L = int(rng.integers(2, 16))
            p = int(rng.integers(0, max(1, n - L)))
            s[p:p + L] = alphabet[int(rng.integers(0, 4))]
        for L in long_runs:  # long homopolymers of chosen lengths (deletion-run stress)
            p = int(rng.integers(0, max(1, n - L)))
            s[p:p + L] = alphabet[int(rng.integers(0, 4))]
        for _ in range(n_runs):
            L = int(rng.integers(1, 40))
            p = int(rng.integers(0, max(1, n - L)))
            s[p:p + L] = ord("N")
        for _ in range(iupac):
            p = int(rng.integers(0, n))
            s[p] = b"RYKMSWBDHV"[int(rng.integers(0, 10))]
        if lowercase_frac > 0:
            k = int(n * lowercase_frac)
            p = int(rng.integers(0, max(1, n - k)))
            seg = s[p:p + k]
            up = (seg >= 65) & (seg <= 90)
            seg[up] += 32
        out.append((name, s.tobytes()))
    return out


def write_fasta(path, contigs, width=70):
    with open(path, "wb") as f:
        for name, s in contigs:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + b"\n")


def read_fasta(path):
    """Restates how get_genome_inf/get_genome_seq see a FASTA (pbsim.cpp:914-965, :1014-1029):
    header = text after '>' (truncated to 128 chars), body = lines concatenated as-is."""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    contigs, name, parts = [], None, []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\n")
            if line.startswith(b">"):
                if name is not None:
                    contigs.append((name, b"".join(parts)))
                name, parts = line[1:129].decode(), []
            else:
                parts.append(line)
    if name is not None:
        contigs.append((name, b"".join(parts)))
    return contigs


def synth_set(seed, n, len_lo=300, len_hi=4000, max_exp=4, lowercase_first=0.0, iupac=0.0, hp_plants=0,
              long_every=0):
    """Deterministic transcript / template set: [(name, plus, minus, bases)].  lowercase_first: share of
    sequences whose text is lower-case (exercises the toupper-from-1 quirk); hp_plants: homopolymers of 9-14."""
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        ln = int(rng.integers(len_lo, len_hi))
        if long_every and t % long_every == long_every - 1:
            ln = int(rng.integers(11000, 14000))  # longer than one fgets buffer (BUF_SIZE 10240)
        s = rng.integers(0, 4, ln).astype(np.uint8)
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[s].copy()
        for _ in range(hp_plants):
            k = int(rng.integers(9, 15))
            p = int(rng.integers(0, max(1, ln - k)))
            s[p:p + k] = s[p]
        if iupac > 0:
            m = rng.random(ln) < iupac
            s[m] = np.frombuffer(b"NRYKM", dtype=np.uint8)[rng.integers(0, 5, int(m.sum()))]
        b = s.tobytes()
        if rng.random() < lowercase_first:
            b = b.lower()
        out.append(("TR%05d.%d" % (t + 1, t % 7), int(rng.integers(0, max_exp + 1)), int(rng.integers(0, max_exp + 1)), b))
    return out


def synth_transcripts(seed, n, n_reads):
    """BASELINE config 4 in small: log-normal lengths (median 1.5 kb), Zipf expression, both strands.
    Returns [(name, plus, minus, bases)] with sum(plus + minus) ~ n_reads."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.exp(rng.normal(np.log(1500.0), 0.75, n)).astype(np.int64), 200, 100000)
    w = 1.0 / np.arange(1, n + 1) ** 0.9
    expr = rng.permutation(np.maximum(1, (w / w.sum() * n_reads)).astype(np.int64))
    out = []
    for t in range(n):
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(lens[t]))].tobytes()
        plus = int((expr[t] + 1) // 2)
        out.append(("T%06d" % (t + 1), plus, int(expr[t] - plus), s))
    return out


def write_transcripts(path, seqset):
    """the 4-column table get_transcript_inf reads (pbsim.cpp:1095-1120): id, plus, minus, sequence"""
    with open(path, "wb") as f:
        for name, plus, minus, s in seqset:
            f.write(name.encode() + b"\t%d\t%d\t" % (plus, minus) + s + b"\n")


def write_templates(path, seqset, width=70):
    write_fasta(path, [(x[0], x[3]) for x in seqset], width)


_SHIM_TEXT = {
    "gzip": '#!/bin/sh\ncat\n[ -n "$PBSIM_SHIM_DONE_DIR" ] && touch "$PBSIM_SHIM_DONE_DIR/done.$$"\nexit 0\n',
    "samtools": '#!/bin/sh\n# samtools view -b -o FILE -\ncat > "$4"\n'
                '[ -n "$PBSIM_SHIM_DONE_DIR" ] && touch "$PBSIM_SHIM_DONE_DIR/done.$$"\nexit 0\n',
}


def ensure_shims():
    """the shims must leave their completion markers (run_reference waits for them): rewrite older ones"""
    if not os.path.isdir(REF_DIR):
        return
    os.makedirs(SHIMS, exist_ok=True)
    for name, text in _SHIM_TEXT.items():
        path = os.path.join(SHIMS, name)
        try:
            with open(path) as f:
                ok = "PBSIM_SHIM_DONE_DIR" in f.read()
        except OSError:
            ok = False
        if not ok:
            with open(path, "w") as f:
                f.write(text)
            os.chmod(path, 0o755)


def run_reference(args, logrand=False, keep_dir=None, real_gzip=False, timeout=3600):
    """Run the reference with `args` (list, without --prefix) in a scratch dir.
    Returns dict: stderr(str), files{name: bytes}, draws(int32 array)|None, marks(int64 array)|None,
    wall(seconds)."""
    ensure_shims()
    work = keep_dir or tempfile.mkdtemp(prefix="pbsim_ref_")
    os.makedirs(work, exist_ok=True)
    done_dir = os.path.join(work, "done")
    os.makedirs(done_dir, exist_ok=True)
    env = dict(os.environ)
    if not real_gzip:
        env["PATH"] = SHIMS + ":" + env.get("PATH", "")
    env["PBSIM_SHIM_DONE_DIR"] = done_dir
    if logrand:
        env["PBSIM_DRAW_LOG"] = os.path.join(work, "draws.bin")
        env["PBSIM_MARK_LOG"] = os.path.join(work, "marks.bin")
    exe = REF_BIN_LOG if logrand else REF_BIN
    t0 = time.time()
    p = subprocess.run([exe] + list(args) + ["--prefix", "out"], cwd=work, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    wall = time.time() - t0
    stderr = p.stderr.decode(errors="replace")
    res = {"stderr": stderr, "returncode": p.returncode, "wall": wall, "files": {}, "draws": None, "marks": None}
    # the reference fclose()s its popen streams and exits without waiting for the children
    # (pbsim.cpp:748-753): wait until every shim has finished writing
    outs = [f for f in glob.glob(os.path.join(work, "out*")) if not f.endswith(".ref")]
    if not real_gzip:
        deadline = time.time() + 60
        while len(os.listdir(done_dir)) < len(outs) and time.time() < deadline:
            time.sleep(0.02)
            outs = [f for f in glob.glob(os.path.join(work, "out*")) if not f.endswith(".ref")]
    for f in sorted(glob.glob(os.path.join(work, "out*"))):
        with open(f, "rb") as fh:
            res["files"][os.path.basename(f)] = fh.read()
    if logrand:
        # (a run that stops before its first rand() call leaves no log)
        dlog, mlog = os.path.join(work, "draws.bin"), os.path.join(work, "marks.bin")
        res["draws"] = np.fromfile(dlog, dtype=np.int32) if os.path.exists(dlog) else np.zeros(0, np.int32)
        res["marks"] = np.fromfile(mlog, dtype=np.int64) if os.path.exists(mlog) else np.zeros(0, np.int64)
    if keep_dir is None:
        shutil.rmtree(work, ignore_errors=True)
    return res


def set_stats_block(stderr):
    """the ':::: Simulation stats ::::' block of a transcript / template run"""
    body = stderr.split(":::: Simulation stats ::::")[1].split(":::: System utilization")[0]
    return ":::: Simulation stats ::::" + body


def split_stats_blocks(stderr):
    """{seq_num: text of ':::: Simulation stats (ref.N) ::::' block}"""
    blocks = {}
    parts = stderr.split(":::: Simulation stats (ref.")
    for part in parts[1:]:
        num = int(part.split(")")[0])
        body = part.split(":::: System utilization")[0]
        blocks[num] = ":::: Simulation stats (ref." + body
    return blocks
