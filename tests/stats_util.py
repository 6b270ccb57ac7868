"""Derive distributions from FASTQ + MAF text (works on reference output, oracle output and engine output alike)."""
import numpy as np


THIN = 1024   # quality values this far apart inside a read are treated as independent draws (the chain mixes in tens
              # of positions): the thinned histogram is what the chi-square test on qualities uses


def sam_to_fastq(sam):
    """the SAM text the reference pipes into `samtools view -b` (pbsim.cpp:2322-2333) as FASTQ records"""
    out = []
    for line in sam.split(b"\n"):
        if not line or line[:1] == b"@":
            continue
        f = line.split(b"\t")
        out.append(b"@" + f[0] + b"\n" + f[9] + b"\n+" + f[0] + b"\n" + f[10] + b"\n")
    return b"".join(out)


def parse_outputs(reads, maf):
    """Returns dict: lengths, accuracy (per read; from qualities), qv_hist[94], qv_thin[94] (every THIN-th position of
    every read), events = (match, sub, ins, del), strand counts.  MAF rows are compared column by column."""
    fq = reads.split(b"\n")
    n = len(fq) // 4
    lengths = np.array([len(fq[4 * k + 1]) for k in range(n)], dtype=np.int64)
    qprob = 10.0 ** (-np.arange(94) / 10.0)
    qv_hist = np.zeros(94, dtype=np.int64)
    qv_thin = np.zeros(94, dtype=np.int64)
    acc = np.zeros(n)
    for k in range(n):
        q = np.frombuffer(fq[4 * k + 3], dtype=np.uint8).astype(np.int64) - 33
        qv_hist += np.bincount(q, minlength=94)[:94]
        qv_thin += np.bincount(q[THIN // 2::THIN], minlength=94)[:94]
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
    return dict(lengths=lengths, accuracy=acc, err_accuracy=err_acc, qv_hist=qv_hist, qv_thin=qv_thin, events=ev,
                per_read=per_read, plus=plus, n=n)


def reduce_for_fixture(st):
    """what a fixture keeps of parse_outputs (a few KB per run)"""
    return dict(lengths=st["lengths"].astype(np.int32), accuracy=st["accuracy"].astype(np.float32),
                err_accuracy=st["err_accuracy"].astype(np.float32), qv_hist=st["qv_hist"], qv_thin=st["qv_thin"],
                events=st["events"], per_read=st["per_read"].astype(np.int32), plus=st["plus"], n=st["n"])


def chi2_error_mix(a, b):
    """chi-square test (3 degrees of freedom) that the pooled substitution / insertion / deletion rates per read base
    are equal in two runs.  Reads are the independent units (a read's accuracy is drawn once, so positions inside a
    read are not): the covariance of the three ratio estimators comes from the reads (delta method), which makes the
    Wald statistic valid where a contingency test on the pooled counts would be wildly over-powered."""
    from scipy import stats

    def est(x):
        ln = np.maximum(1, x["lengths"]).astype(np.float64)
        c = x["per_read"][:, 1:4].astype(np.float64)
        tot = ln.sum()
        p = c.sum(0) / tot
        resid = (c - ln[:, None] * p[None, :]) / tot          # linearised ratio estimator
        cov = resid.T @ resid * (len(ln) / max(1, len(ln) - 1))
        return p, cov

    pa, ca = est(a)
    pb, cb = est(b)
    d = pa - pb
    chi2 = float(d @ np.linalg.pinv(ca + cb) @ d)
    return chi2, float(stats.chi2.sf(chi2, 3))


def chi2_qualities(a, b, min_expected=8):
    """chi-square two-sample test on the THINNED quality histograms (positions THIN apart: independent draws)"""
    from scipy import stats
    ha, hb = a["qv_thin"].astype(np.float64), b["qv_thin"].astype(np.float64)
    keep = (ha + hb) >= 2 * min_expected
    ta = np.append(ha[keep], ha[~keep].sum())
    tb = np.append(hb[keep], hb[~keep].sum())
    if ta[-1] + tb[-1] == 0:
        ta, tb = ta[:-1], tb[:-1]
    if len(ta) < 2:
        return 0.0, 1.0
    chi2, p, dof, _ = stats.chi2_contingency(np.vstack([ta, tb]))
    return float(chi2), float(p)


def compare(a, b, method, alpha=1e-3, chi2=True):
    """Two-sample comparison of engine/oracle output `a` with the reference fixture `b`.
    Stated tolerances: every KS test p > alpha (reads are i.i.d. units, so read-level tests are exact);
    pooled error rates within 3 % relative; quality histogram total variation < 0.01; and, where both sides carry
    the thinned quality histogram, the chi-square tests of chi2_error_mix / chi2_qualities at the same alpha."""
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
    if chi2 and "qv_thin" in a and "qv_thin" in b:
        res["chi2_error_mix"], res["p_error_mix"] = chi2_error_mix(a, b)
        if method == "qshmm":
            res["chi2_qv"], res["p_qv"] = chi2_qualities(a, b)
    # the fixed 3 % bound on the pooled rates is the fall-back where the (noise-aware) chi-square test is not available
    rel_ok = "p_error_mix" in res or all(v < 0.03 for k, v in res.items() if k.startswith("rel_"))
    ok = all(v > alpha for k, v in res.items() if k.startswith("ks_") or k.startswith("p_")) and rel_ok and \
        res["tv_qv"] < 0.01
    return ok, res
