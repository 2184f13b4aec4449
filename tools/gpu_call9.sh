#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
UVD="python bench.py --workload uvd --no-e2e --no-cpu-baseline --no-separate --steps 40"
i=0
for v in base rpl2 rpl4 map2 base rpl2 rpl4 map2; do
  i=$((i+1))
  if [ $v = base ]; then timeout 200 $UVD > $OUT/c9_${i}_$v.json 2>> $OUT/c9_bench.err
  else PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_$v.so timeout 200 $UVD > $OUT/c9_${i}_$v.json 2>> $OUT/c9_bench.err; fi
done
for f in $OUT/c9_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels']], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR')
PY
done
