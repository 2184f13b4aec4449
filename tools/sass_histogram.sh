#!/usr/bin/env bash
# SASS opcode histogram of the shipped objects: proof that the hot path is tcgen05 / TMA / TMEM code (B200_PROFILING.md
# lists the mnemonics).  Runs in the CPU container:  bash tools/sass_histogram.sh > profiles/r02_sass_histogram.txt
set -u
cd "$(dirname "$0")/.."
python -c 'from psgd_tf_b200 import build; build.build()' > /dev/null 2>&1
for obj in gemm_tc uvd kron_stream elementwise splu dense linalg comm; do
  echo "== psgd_tf_b200/_C/$obj.o"
  cuobjdump -sass psgd_tf_b200/_C/$obj.o | grep -E '^\s+/\*[0-9a-f]+\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed -E 's/;$//' \
    | grep -E '^(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCATOM|SYNCS|FFMA2|FFMA|LDGSTS|STG|LDG|LDS|STS|ELECT|UCGABAR|ACQBULK|BAR|ATOM|RED|MUFU|DFMA|HMMA)' \
    | sed -E 's/^(UTCHMMA|UTCBAR|UTMALDG|UBLKCP|LDTM|STTM|SYNCS|FFMA2|LDGSTS|UCGABAR)(\.[A-Z0-9_.]+)?$/\1\2/' | sort | uniq -c | sort -rn | head -40
done
echo "== per kernel (gemm_tc.o): instructions / UTCHMMA / UTMALDG / LDTM / STTM"
cuobjdump -sass psgd_tf_b200/_C/gemm_tc.o | awk '/Function : /{name=$3} /^[ \t]+\/\*[0-9a-f]+\*\//{n[name]++; if ($0 ~ /UTCHMMA/) a[name]++; if ($0 ~ /UTMALDG/) b[name]++; if ($0 ~ /LDTM/) c[name]++; if ($0 ~ /STTM/) d[name]++} END{for (k in n) printf "%-90s %6d %4d %4d %4d %4d\n", k, n[k], a[k], b[k], c[k], d[k]}' | sort
