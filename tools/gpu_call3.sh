#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_splu.py tests/test_gpu_demos.py -x -q -m gpu -k "norm_scale or dense or kron_all or golden or splu or demo" > $OUT/c3_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c3_status.txt
timeout 300 python bench_aux.py > $OUT/c3_aux.jsonl 2> $OUT/c3_bench.err; echo "aux rc=$?" >> $OUT/c3_status.txt
UVD="python bench.py --workload uvd --no-e2e --no-cpu-baseline --no-separate --steps 40"
timeout 200 $UVD > $OUT/c3_uvd_base.json 2>> $OUT/c3_bench.err; echo "base rc=$?" >> $OUT/c3_status.txt
for v in rcp r2w4 r2w4rcp r4w3 r4w3rcp; do
  PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_$v.so timeout 200 $UVD > $OUT/c3_uvd_$v.json 2>> $OUT/c3_bench.err; echo "$v rc=$?" >> $OUT/c3_status.txt
done
PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_r4w3rcp.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uvd" > $OUT/c3_pytest_r4w3rcp.log 2>&1; echo "pytest-variant rc=$?" >> $OUT/c3_status.txt
AUX='ns_stats|col_wsum|col_finish|row_dot|max_kernel|small._kernel|pass._kernel'
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"$AUX" --csv --log-file $OUT/r01d_aux_launches.csv python bench_aux.py > $OUT/c3_aux_launches.log 2>&1; echo "aux-launch rc=$?" >> $OUT/c3_status.txt
cat $OUT/c3_status.txt; tail -3 $OUT/c3_pytest.log; cat $OUT/c3_aux.jsonl | cut -c1-260
for f in $OUT/c3_uvd_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels']])
PY
done
