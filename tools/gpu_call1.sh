#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, default bench, UVd profile, ncu data for the aux streaming kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/c1_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uvd" > $OUT/c1_pytest_uvd.log 2>&1; echo "pytest-uvd rc=$?" >> $OUT/c1_status.txt
timeout 900 python -m pytest tests -q -m gpu -k "not uvd or class" > $OUT/c1_pytest_rest.log 2>&1; echo "pytest-rest rc=$?" >> $OUT/c1_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/c1_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/c1_status.txt
timeout 600 python bench.py > $OUT/c1_bench.json 2> $OUT/c1_bench.err; echo "bench rc=$?" >> $OUT/c1_status.txt
timeout 200 python bench.py --workload uvd --uvd-form separate-3sweep --no-e2e --no-cpu-baseline > $OUT/c1_bench_3sweep.json 2>> $OUT/c1_bench.err; echo "bench3 rc=$?" >> $OUT/c1_status.txt
timeout 500 bash tools/profile_uvd.sh r01b > $OUT/c1_profile.log 2>&1; echo "profile rc=$?" >> $OUT/c1_status.txt
AUX='ns_stats|ns_apply|ns_finish|col_wsum|col_finish|row_dot|norm_new|scale_new|balance_kernel|rescale|diag_|xmat_'
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"$AUX" --csv --log-file $OUT/r01b_aux_launches.csv python bench_aux.py > $OUT/c1_aux_launches.log 2>&1; echo "aux-launch rc=$?" >> $OUT/c1_status.txt
timeout 300 ncu --clock-control none --set full --import-source on -k regex:'ns_stats_kernel|col_wsum_kernel|row_dot_kernel|ns_apply_kernel' -s 100 -c 24 -f -o /tmp/aux_full python bench_aux.py > $OUT/c1_aux_full.log 2>&1
ncu -i /tmp/aux_full.ncu-rep --page raw --csv > $OUT/r01b_aux_full_raw.csv 2>/dev/null
ncu -i /tmp/aux_full.ncu-rep --page source --csv --kernel-name regex:'ns_stats_kernel' --launch-count 1 > $OUT/r01b_aux_full_source.csv 2>/dev/null
echo "aux-full done" >> $OUT/c1_status.txt
cat $OUT/c1_status.txt
tail -3 $OUT/c1_pytest_uvd.log $OUT/c1_pytest_rest.log
head -c 1500 $OUT/c1_bench.json
