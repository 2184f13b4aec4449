#!/usr/bin/env bash
# ncu --set full capture of ONE dense 4096^3 product through the tcgen05 3xTF32 engine (tools/gemm_debug.py perf),
# exported on the box to small CSVs (raw metrics + per-instruction source page).   gpurun -- 'bash tools/profile_gemm.sh TAG'
set -u
TAG=${1:-gemm}
OUT=gpurun_out; TMP=/tmp/psgd_prof; mkdir -p $OUT $TMP
ncu --clock-control none --set full --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $TMP/$TAG python tools/gemm_debug.py perf > $OUT/${TAG}.log 2>&1
ncu -i $TMP/$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $TMP/$TAG.ncu-rep --page source --csv > $OUT/${TAG}_source.csv 2>/dev/null
rm -f $TMP/$TAG.ncu-rep
ls -la $OUT | tail -5
