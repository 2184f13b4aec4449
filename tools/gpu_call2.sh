#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "norm_scale or dense or kron_all or golden" > $OUT/c2_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c2_status.txt
timeout 400 python bench.py --workload uvd > $OUT/c2_bench_uvd.json 2> $OUT/c2_bench.err; echo "bench-uvd rc=$?" >> $OUT/c2_status.txt
timeout 300 python bench_aux.py > $OUT/c2_aux.jsonl 2>> $OUT/c2_bench.err; echo "aux rc=$?" >> $OUT/c2_status.txt
AUX='ns_stats|ns_apply|ns_finish|col_wsum|col_finish|row_dot|norm_new|scale_new|balance_kernel|rescale|splu'
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"$AUX" --csv --log-file $OUT/r01c_aux_launches.csv python bench_aux.py > $OUT/c2_aux_launches.log 2>&1; echo "aux-launch rc=$?" >> $OUT/c2_status.txt
cat $OUT/c2_status.txt; tail -3 $OUT/c2_pytest.log; head -c 600 $OUT/c2_bench_uvd.json; echo; cat $OUT/c2_aux.jsonl
