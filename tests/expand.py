"""TEST CODE — reference expansion of the engine's pass-1 event streams into text records.

This is the executable specification of pass 2 (pbsim_b200/csrc/emit.cuh): given the genome, a read
plan and the event stream of one (read, pass), produce the read, qualities and the two MAF rows,
then format FASTQ / SAM / MAF exactly as the reference does (pbsim.cpp:2318-2383).  numpy only.
"""
import numpy as np

_COMP = np.arange(256, dtype=np.uint8)
for a, b in ((b"A", b"T"), (b"T", b"A"), (b"G", b"C"), (b"C", b"G")):
    _COMP[a[0]] = b[0]

# substitution alphabets (pbsim.cpp:5481-5486), indexed by the reference base
_SUB = np.zeros((256, 3), dtype=np.uint8)
_SUB[ord("A")] = list(b"TGC")
_SUB[ord("T")] = list(b"AGC")
_SUB[ord("G")] = list(b"ATC")
_SUB[ord("C")] = list(b"ATG")
_NT4 = np.frombuffer(b"ATGC", dtype=np.uint8)
_IS_ACGT = np.zeros(256, dtype=bool)
for ch in b"ACGT":
    _IS_ACGT[ch] = True


def window(genome_upper, offset, wlen, minus):
    w = np.frombuffer(genome_upper, dtype=np.uint8, count=wlen, offset=offset)
    if minus:
        w = _COMP[w[::-1]]
    return w


def _read_bases(kind, info, nt):
    rb = nt.copy()
    sub = kind == 1
    acgt = _IS_ACGT[nt]
    s1 = sub & acgt
    rb[s1] = _SUB[nt[s1], np.minimum(info[s1], 2)]
    s2 = sub & ~acgt
    rb[s2] = _NT4[info[s2] & 3]
    ins = (kind == 2) & (info < 4)
    rb[ins] = _NT4[info[ins] & 3]
    return rb


def expand_qshmm(ev, genome_upper, offset, wlen, minus, partial=False):
    """ev: uint16 entries.  Returns (seq, qual, maf_ref, maf_read) as uint8 arrays.
    partial: --method sample, where a read may end before its window is used up"""
    ev = np.asarray(ev, dtype=np.uint16).astype(np.int64)
    kind = (ev >> 7) & 3
    cont = kind == 3
    qv = ev & 0x7F
    info = (ev >> 9) & 7
    nd = np.where(cont, (ev & 0x7F) | ((ev >> 9) << 7), (ev >> 12) & 15)
    is_base = ~cont
    adv = (is_base & (kind != 2)).astype(np.int64)
    ref_adv = adv + nd
    col_adv = is_base.astype(np.int64) + nd
    R = np.cumsum(ref_adv) - ref_adv
    Ccol = np.cumsum(col_adv) - col_adv
    ncol = int(col_adv.sum())
    assert int(ref_adv.sum()) == wlen or (partial and int(ref_adv.sum()) < wlen), (int(ref_adv.sum()), wlen)
    W = window(genome_upper, offset, wlen, minus)
    b = np.nonzero(is_base)[0]
    nt = W[R[b]]
    rb = _read_bases(kind[b], info[b], nt)
    seq = rb
    qual = (qv[b] + 33).astype(np.uint8)
    maf_ref = np.empty(ncol, dtype=np.uint8)
    maf_read = np.empty(ncol, dtype=np.uint8)
    maf_read[Ccol[b]] = rb
    maf_ref[Ccol[b]] = np.where(kind[b] == 2, ord("-"), nt)
    # deletion columns
    d = np.nonzero(nd > 0)[0]
    if len(d):
        reps = nd[d]
        first_col = np.repeat(Ccol[d] + is_base[d], reps)
        first_ref = np.repeat(R[d] + adv[d], reps)
        j = np.arange(int(reps.sum())) - np.repeat(np.cumsum(reps) - reps, reps)
        maf_read[first_col + j] = ord("-")
        maf_ref[first_col + j] = W[first_ref + j]
    if minus:
        maf_ref = _COMP[maf_ref[::-1]]
        maf_read = _COMP[maf_read[::-1]]
    return seq, qual, maf_ref, maf_read


def expand_errhmm(ev, genome_upper, offset, wlen, minus):
    ev = np.asarray(ev, dtype=np.uint8).astype(np.int64) & 0x1F  # bits 5-7: pass-1 bookkeeping of the segment path
    kind = ev & 3
    info = (ev >> 2) & 7
    ref_adv = (kind != 2).astype(np.int64)
    R = np.cumsum(ref_adv) - ref_adv
    assert int(ref_adv.sum()) == wlen
    W = window(genome_upper, offset, wlen, minus)
    nt = W[np.minimum(R, wlen - 1)]
    rb = _read_bases(kind, info, nt)
    maf_read = np.where(kind == 3, ord("-"), rb).astype(np.uint8)
    maf_ref = np.where(kind == 2, ord("-"), nt).astype(np.uint8)
    seq = rb[kind != 3]
    qual = np.full(len(seq), ord("!"), dtype=np.uint8)
    if minus:
        maf_ref = _COMP[maf_ref[::-1]]
        maf_read = _COMP[maf_read[::-1]]
    return seq, qual, maf_ref, maf_read


def count_digit(n):
    return len(str(int(n)))


def format_records(pass_num, id_prefix, seq_num, read_id, pas, offset, wlen, glen, minus, seq, qual, maf_ref,
                   maf_read, accuracy_mean):
    """Returns (reads_bytes, maf_bytes) for one (read, pass)."""
    seq_b, qual_b = seq.tobytes(), qual.tobytes()
    rlen = len(seq_b)
    if pass_num == 1:
        rid = ("%s%d_%d" % (id_prefix, seq_num, read_id)).encode()
        reads = b"@" + rid + b"\n" + seq_b + b"\n+" + rid + b"\n" + qual_b + b"\n"
    else:
        rid = ("%s%d/%d/%d" % (id_prefix, seq_num, read_id, pas)).encode()
        reads = (rid + b"\t4\t*\t0\t255\t*\t*\t0\t0\t" + seq_b + b"\t" + qual_b + b"\tcx:i:3\tip:B:C" + b",9" * rlen
                 + b"\tnp:i:1\tpw:B:C" + b",9" * rlen
                 + ("\tqs:i:0\tqe:i:%d\trq:f:%f\tsn:B:f,10.0,10.0,10.0,10.0\tzm:i:%d\tRG:Z:ffffffff\n"
                    % (rlen - 1, accuracy_mean, read_id)).encode())
    d1 = [3, count_digit(offset), count_digit(wlen), count_digit(glen)]
    d2 = [1 + count_digit(read_id), 1, count_digit(rlen), count_digit(rlen)]
    dn = [max(a, b) for a, b in zip(d1, d2)]
    sp = lambda n: b" " * n  # noqa: E731
    maf = (b"a\ns ref" + sp(dn[0] - d1[0]) + sp(dn[1] - d1[1]) + b" %d" % offset + sp(dn[2] - d1[2])
           + b" %d +" % wlen + sp(dn[3] - d1[3]) + b" %d " % glen + maf_ref.tobytes() + b"\n"
           + b"s " + rid + sp(dn[0] - d2[0]) + sp(dn[1] - d2[1]) + b" 0" + sp(dn[2] - d2[2])
           + b" %d %s" % (rlen, b"-" if minus else b"+") + sp(dn[3] - d2[3]) + b" %d " % rlen
           + maf_read.tobytes() + b"\n\n")
    return reads, maf
