#!/bin/bash
# Wall clock of the whole tool (parse FASTA, simulate, write .fq.gz/.maf.gz) on a synthetic 200 Mbp, 4-contig genome:
# the B200 driver against the unmodified reference (real gzip children), the latter on a 1/20 sample.
set -e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
W=$(mktemp -d)
python - "$W" <<'PY'
import sys, numpy as np
w = sys.argv[1]
rng = np.random.default_rng(7)
def fasta(path, sizes):
    with open(path, "wb") as f:
        for i, n in enumerate(sizes, 1):
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
            f.write(b">chr%d\n" % i)
            rows = s[: n // 70 * 70].reshape(-1, 70)
            f.write(b"\n".join(r.tobytes() for r in rows) + b"\n")
fasta(w + "/g200.fa", [80000000, 60000000, 40000000, 20000000])
fasta(w + "/g10.fa", [4000000, 3000000, 2000000, 1000000])
PY
MODEL=$(python -c "from tests.golden_util import model_path; print(model_path('QSHMM-RSII.model'))")
cd "$W"
echo "== B200 driver, 200 Mbp x depth 20 (4.0 Gbase), qshmm RSII"
T0=$(date +%s.%N)
"$OLDPWD/pbsim_b200/bin/pbsim" --strategy wgs --method qshmm --qshmm "$MODEL" \
  --genome g200.fa --depth 20 --seed 1 --prefix b200 2> b200.err || { tail -5 b200.err; exit 1; }
T1=$(date +%s.%N)
echo "wall $(python -c "print(round($T1-$T0,2))") s"
tail -4 b200.err
ls -la b200_0001.fq.gz b200_0001.maf.gz | awk '{print $5, $9}'
zcat b200_0004.fq.gz | head -2 | cut -c1-80
if [ -x "$OLDPWD/oracle/_ref/pbsim" ]; then
  echo "== reference, 10 Mbp x depth 20 (0.2 Gbase), real gzip children"
  T0=$(date +%s.%N)
  "$OLDPWD/oracle/_ref/pbsim" --strategy wgs --method qshmm --qshmm "$MODEL" \
    --genome g10.fa --depth 20 --seed 1 --prefix ref 2> ref.err || true
  T1=$(date +%s.%N)
  echo "wall $(python -c "print(round($T1-$T0,2))") s (the gzip children may still be flushing)"
  tail -3 ref.err
fi
rm -rf "$W"
