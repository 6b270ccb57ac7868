#!/bin/bash
# Wall clock of the whole tool on a human-sized genome (round 2): parse a 3.1 Gbp, 24-contig FASTA, simulate ONT-like
# 50 kb reads (qshmm, QSHMM-ONT), write .ref / .fq.gz / .maf.gz (gzip members written by the GPU) to a RAM disk.
# usage: profiles/cli_wallclock_r02.sh [depth]   (default: what fits the RAM disk, at most 10)
set -e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
W=/dev/shm/pbsim_wall_$$
mkdir -p "$W" 2>/dev/null || W=$(mktemp -d)
trap 'rm -rf "$W"' EXIT
T0=$(date +%s.%N)
python - "$W" <<'PY'
import sys, numpy as np
w = sys.argv[1]
mbp = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
rng = np.random.default_rng(7)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
with open(w + "/genome.fa", "wb") as f:
    for i, m in enumerate(mbp, 1):
        n = m * 1000000 // 70 * 70
        rows = np.empty((n // 70, 71), dtype=np.uint8)
        rows[:, :70] = lut[rng.integers(0, 4, n, dtype=np.uint8)].reshape(-1, 70)
        rows[:, 70] = 10
        f.write(b">chr%d synthetic\n" % i)
        f.write(rows.tobytes())
PY
T1=$(date +%s.%N)
echo "genome.fa: $(stat -c %s "$W/genome.fa") bytes, written in $(python -c "print(round($T1-$T0,1))") s"
AVAIL_GB=$(df -BG --output=avail "$W" | tail -1 | tr -dc 0-9)
DEPTH=${1:-$(python -c "print(max(1, min(10, int(($AVAIL_GB - 8) / (3.085 * 1.7)))))")}
echo "RAM disk: ${AVAIL_GB} GB free; depth $DEPTH; host cores: $(nproc)"
MODEL=$(python -c "from tests.golden_util import model_path; print(model_path('QSHMM-ONT.model'))")
cd "$W"
# Two runs: on a freshly leased VM the first touch of every page of guest memory is served by the hypervisor, and the
# 50 GB of output pages of the first run pay for that; the second run writes into pages the guest already owns.
# Third run: the 48 output files are symbolic links to /dev/null (the tool without the file system behind it).
for RUN in 1 2 3; do
rm -f b200_*
if [ $RUN = 3 ]; then
  for i in $(seq -f %04g 1 24); do ln -s /dev/null b200_$i.fq.gz; ln -s /dev/null b200_$i.maf.gz; done
fi
T0=$(date +%s.%N)
"$ROOT/pbsim_b200/bin/pbsim" --strategy wgs --method qshmm --qshmm "$MODEL" --genome genome.fa --depth "$DEPTH" \
  --length-mean 50000 --length-sd 35000 --length-max 1000000 --difference-ratio 39:24:36 --seed 1 --prefix b200 \
  2> b200.err || { tail -5 b200.err; exit 1; }
T1=$(date +%s.%N)
BASES=$(python - <<'PY'
import re
t = open("b200.err").read()
n = [int(x) for x in re.findall(r"read num\. : (\d+)", t)]
m = [float(x) for x in re.findall(r"read length mean \(SD\) : ([0-9.]+)", t)]
print(int(sum(a * b for a, b in zip(n, m))))
PY
)
python -c "w=$T1-$T0; print('run $RUN, whole tool: %.2f s wall for %.2f Gbase = %.2f Gbp/s' % (w, $BASES/1e9, $BASES/1e9/w))"
tail -8 b200.err | grep -v "^$"
done
rm -f b200_*.gz
# what the RAM disk takes from 16 plain writers (dd, 4 MiB blocks, 2 GiB each)
T0=$(date +%s.%N)
for i in $(seq 1 16); do dd if=/dev/zero of=dd_$i bs=4M count=512 2>/dev/null & done
wait
T1=$(date +%s.%N)
python -c "print('RAM disk ceiling: 16 dd writers, 34.4 GB in %.2f s = %.2f GB/s' % ($T1-$T0, 34.36/($T1-$T0)))"
rm -f dd_*
du -sh --apparent-size . | cut -f1 | xargs echo "output + input bytes on the RAM disk:"
ls -la b200_0001.ref | awk '{print $5, $9}'
