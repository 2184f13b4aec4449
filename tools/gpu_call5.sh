#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
ncu --clock-control none --set full --import-source on -k regex:'pass._kernel' -s 7 -c 4 -f -o /tmp/splu_full python tools/splu_probe.py 5e7 2 > $OUT/c5_splu_ncu.log 2>&1
ncu -i /tmp/splu_full.ncu-rep --page raw --csv > $OUT/r01d_splu_full_raw.csv 2>/dev/null
for k in pass2 pass4; do
ncu -i /tmp/splu_full.ncu-rep --page source --csv --kernel-name regex:${k}_kernel > $OUT/r01d_splu_${k}_source.csv 2>$OUT/c5_err_$k.txt
ncu -i /tmp/splu_full.ncu-rep --page details --kernel-name regex:${k}_kernel > $OUT/r01d_splu_${k}_details.txt 2>>$OUT/c5_err_$k.txt
done
ls -la $OUT | grep splu
