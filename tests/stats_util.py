"""Derive distributions from FASTQ + MAF text (works on reference output, oracle output and engine output alike)."""
import numpy as np


def parse_outputs(reads, maf):
    """Returns dict: lengths, accuracy (per read; from qualities), qv_hist[94], events = (match, sub, ins, del),
    strand counts.  MAF rows are compared column by column."""
    fq = reads.split(b"\n")
    n = len(fq) // 4
    lengths = np.array([len(fq[4 * k + 1]) for k in range(n)], dtype=np.int64)
    qprob = 10.0 ** (-np.arange(94) / 10.0)
    qv_hist = np.zeros(94, dtype=np.int64)
    acc = np.zeros(n)
    for k in range(n):
        q = np.frombuffer(fq[4 * k + 3], dtype=np.uint8).astype(np.int64) - 33
        qv_hist += np.bincount(q, minlength=94)[:94]
        acc[k] = 1.0 - qprob[q].mean() if len(q) else 0.0
    ev = np.zeros(4, dtype=np.int64)
    per_read = np.zeros((n, 4), dtype=np.int64)
    plus = 0
    err_acc = np.zeros(n)
    for k, blk in enumerate(maf.split(b"\n\n")[:-1]):
        l1, l2 = blk.split(b"\n")[1:3]
        r = np.frombuffer(l1.split()[6], dtype=np.uint8)
        q = np.frombuffer(l2.split()[6], dtype=np.uint8)
        ins = r == ord("-")
        dele = q == ord("-")
        sub = (~ins) & (~dele) & (r != q)
        cnt = np.array([np.count_nonzero((~ins) & (~dele) & (r == q)), np.count_nonzero(sub), np.count_nonzero(ins),
                        np.count_nonzero(dele)])
        ev += cnt
        per_read[k] = cnt
        plus += l2.split()[4] == b"+"
        rl = cnt[0] + cnt[1] + cnt[2]
        err_acc[k] = 1.0 - (cnt[1] + cnt[2] + cnt[3]) / max(1, rl)
    return dict(lengths=lengths, accuracy=acc, err_accuracy=err_acc, qv_hist=qv_hist, events=ev, per_read=per_read,
                plus=plus, n=n)


def compare(a, b, method, alpha=1e-3):
    """Two-sample comparison of engine/oracle output `a` with the reference fixture `b`.
    Stated tolerances: every KS test p > alpha (reads are i.i.d. units, so read-level tests are exact);
    pooled error rates within 3 % relative; quality histogram total variation < 0.01."""
    from scipy import stats
    res = {}
    res["ks_length"] = stats.ks_2samp(a["lengths"], b["lengths"]).pvalue
    key = "accuracy" if method == "qshmm" else "err_accuracy"
    res["ks_accuracy"] = stats.ks_2samp(a[key], b[key]).pvalue
    for j, nm in ((1, "sub"), (2, "ins"), (3, "del")):
        ra = a["per_read"][:, j] / np.maximum(1, a["lengths"])
        rb = b["per_read"][:, j] / np.maximum(1, b["lengths"])
        res["ks_%s_rate" % nm] = stats.ks_2samp(ra, rb).pvalue
        pa = a["events"][j] / a["lengths"].sum()
        pb = b["events"][j] / b["lengths"].sum()
        res["rel_%s" % nm] = abs(pa - pb) / pb
    ha = a["qv_hist"] / a["qv_hist"].sum()
    hb = b["qv_hist"] / b["qv_hist"].sum()
    res["tv_qv"] = 0.5 * np.abs(ha - hb).sum()
    ok = all(v > alpha for k, v in res.items() if k.startswith("ks_")) and \
        all(v < 0.03 for k, v in res.items() if k.startswith("rel_")) and res["tv_qv"] < 0.01
    return ok, res
