#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_splu.py tests/test_gpu_uvd_class.py tests/test_gpu_demos.py -x -q -m gpu > $OUT/c6_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c6_status.txt
python tools/splu_probe.py 5e7 4 > $OUT/c6_splu_time.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:'ns_stats_kernel|col_wsum_kernel' -s 2 -c 2 -f -o /tmp/ns_full python tools/ns_probe.py > $OUT/c6_ns_ncu.log 2>&1
ncu -i /tmp/ns_full.ncu-rep --page raw --csv > $OUT/r01d_ns_full_raw.csv 2>/dev/null
ncu -i /tmp/ns_full.ncu-rep --page source --csv --kernel-name regex:ns_stats_kernel > $OUT/r01d_ns_stats_source.csv 2>/dev/null
ncu -i /tmp/ns_full.ncu-rep --page details --kernel-name regex:ns_stats_kernel > $OUT/r01d_ns_stats_details.txt 2>/dev/null
UVD="python bench.py --workload uvd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-separate"
ncu --clock-control none --set full --import-source on -k regex:'gram_sweep' -s 3 -c 1 -f -o /tmp/gram_full $UVD > $OUT/c6_gram_ncu.log 2>&1
ncu -i /tmp/gram_full.ncu-rep --page source --csv > $OUT/r01d_gram_source.csv 2>/dev/null
ncu -i /tmp/gram_full.ncu-rep --page details > $OUT/r01d_gram_details.txt 2>/dev/null
cat $OUT/c6_status.txt; tail -3 $OUT/c6_pytest.log; cat $OUT/c6_splu_time.log
