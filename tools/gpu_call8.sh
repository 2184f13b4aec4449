#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uvd" > $OUT/c8_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c8_status.txt
UVD="python bench.py --workload uvd --no-e2e --no-cpu-baseline --steps 40"
timeout 200 $UVD > $OUT/c8_uvd_base.json 2>> $OUT/c8_bench.err
for v in rpl6 rpl8 map2 map2rpl6; do
  PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_$v.so timeout 200 $UVD > $OUT/c8_uvd_$v.json 2>> $OUT/c8_bench.err; echo "$v rc=$?" >> $OUT/c8_status.txt
done
PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_map2.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uvd" > $OUT/c8_pytest_map2.log 2>&1; echo "pytest-map2 rc=$?" >> $OUT/c8_status.txt
cat $OUT/c8_status.txt; tail -2 $OUT/c8_pytest.log
for f in $OUT/c8_uvd_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels']], d['separate_calls']['value'])
PY
done
