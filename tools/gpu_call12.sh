#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kron_step_tail.py tests/test_gpu_demos.py -x -q -m gpu -k "kron or norm or demo or golden" > $OUT/c12_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c12_status.txt
timeout 300 python bench_aux.py > $OUT/c12_aux.jsonl 2> $OUT/c12_bench.err; echo "aux rc=$?" >> $OUT/c12_status.txt
cat $OUT/c12_status.txt; tail -3 $OUT/c12_pytest.log; cut -c1-250 $OUT/c12_aux.jsonl
