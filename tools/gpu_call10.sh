#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_splu.py tests/test_gpu_uvd_class.py -x -q -m gpu -k "uvd or splu or class or step" > $OUT/c10_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c10_status.txt
python tools/splu_probe.py 5e7 4 > $OUT/c10_splu_time.log 2>&1
timeout 300 python bench.py --workload uvd --no-cpu-baseline --steps 40 > $OUT/c10_uvd.json 2> $OUT/c10_bench.err
cat $OUT/c10_status.txt; tail -2 $OUT/c10_pytest.log; cat $OUT/c10_splu_time.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/c10_uvd.json')); print(d['value'], d['ms_per_step'], [(k['kernel'],k['avg_ms'],k['frac']) for k in d['kernels']], d['separate_calls'], d['e2e']['value'], d['clocks'])
PY
