#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
UVD="python bench.py --workload uvd --no-e2e --no-cpu-baseline --no-separate --steps 40"
timeout 200 $UVD > $OUT/c7_uvd_base.json 2>> $OUT/c7_bench.err
for v in deep rpl4 rpl4deep rpl1deep; do
  PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_$v.so timeout 200 $UVD > $OUT/c7_uvd_$v.json 2>> $OUT/c7_bench.err; echo "$v rc=$?" >> $OUT/c7_status.txt
done
PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_deep.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uvd" > $OUT/c7_pytest_deep.log 2>&1; echo "pytest-deep rc=$?" >> $OUT/c7_status.txt
ncu --clock-control none --set full --import-source on -k regex:'pass._kernel' -s 7 -c 4 -f -o /tmp/splu_full python tools/splu_probe.py 5e7 2 > $OUT/c7_splu_ncu.log 2>&1
ncu -i /tmp/splu_full.ncu-rep --page raw --csv > $OUT/r01e_splu_full_raw.csv 2>/dev/null
for k in pass2 pass3 pass4; do
ncu -i /tmp/splu_full.ncu-rep --page source --csv --kernel-name regex:${k}_kernel > $OUT/r01e_splu_${k}_source.csv 2>/dev/null
done
cat $OUT/c7_status.txt
for f in $OUT/c7_uvd_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels']])
PY
done
