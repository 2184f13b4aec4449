#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python tools/splu_probe.py 5e7 3 > $OUT/c4_splu_time.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:'pass._kernel' -s 5 -c 4 -f -o /tmp/splu_full python tools/splu_probe.py 5e7 2 > $OUT/c4_splu_ncu.log 2>&1
ncu -i /tmp/splu_full.ncu-rep --page raw --csv > $OUT/r01d_splu_full_raw.csv 2>/dev/null
ncu -i /tmp/splu_full.ncu-rep --page source --csv --kernel-name regex:'pass4_kernel' --launch-count 1 > $OUT/r01d_splu_pass4_source.csv 2>/dev/null
ncu -i /tmp/splu_full.ncu-rep --page details --kernel-name regex:'pass4_kernel' --launch-count 1 > $OUT/r01d_splu_pass4_details.txt 2>/dev/null
cat $OUT/c4_splu_time.log
