#!/usr/bin/env bash
# ncu evidence for the two headline workloads (run under gpurun on ONE B200; see B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/profile_gpu.sh r01'
# Writes only small CSV/text files to gpurun_out/ (the .ncu-rep files are exported to CSV on the box and deleted:
# gpurun_out/ is capped at 64 MiB).  tools/summarise_profiles.py turns them into profiles/<tag>_*.
set -u
TAG=${1:-r01}
OUT=gpurun_out
TMP=/tmp/psgd_prof
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
UVD="python bench.py --workload uvd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
KRON="python bench.py --workload kron --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
OURS='gram_sweep|map_sweep|reduce_partials|uvd_small|maxabs2|balance|zero_small|exchange|gemm_tc|gemm_simt|tri_inv|rescale|trsm'

# 1. launch lists: every launch of OUR kernels in the timed region with its device time (single pass)
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 33 --csv --log-file $OUT/${TAG}_uvd_launches.csv $UVD > $OUT/${TAG}_uvd_launches.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 768 -c 256 --csv --log-file $OUT/${TAG}_kron_launches.csv $KRON > $OUT/${TAG}_kron_launches.log 2>&1

# 2. full captures of the dominant kernels -> raw metrics CSV (+ per-instruction source page for the top kernel)
export_rep() {   # $1 = rep basename, $2 = kernel regex for the source page
  ncu -i $TMP/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $TMP/$1.ncu-rep --page source --csv --kernel-name regex:"$2" --launch-count 1 > $OUT/$1_source.csv 2>/dev/null
  rm -f $TMP/$1.ncu-rep
}
$NCU --set full --import-source on -k regex:'gram_sweep|map_sweep' -s 15 -c 5 -f -o $TMP/${TAG}_uvd_full $UVD > $OUT/${TAG}_uvd_full.log 2>&1
export_rep ${TAG}_uvd_full gram_sweep
# Kron stack with 6 layers per grouped launch (same kernels, same tile shapes; keeps ncu's save/restore small):
# launches 0-1 of a step are the two big products forming A, the last 8 are grad1/Ql'/grad2/Qr' and the 4 apply GEMMs
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 408 -c 2 -f -o $TMP/${TAG}_kron_full_head $KRON --layers 6 > $OUT/${TAG}_kron_full_head.log 2>&1
export_rep ${TAG}_kron_full_head gemm_tc
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 536 -c 8 -f -o $TMP/${TAG}_kron_full_tail $KRON --layers 6 > $OUT/${TAG}_kron_full_tail.log 2>&1
export_rep ${TAG}_kron_full_tail gemm_tc
# the GEMM engine alone: one dense 4096^3 product (no triangular hints)
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $TMP/${TAG}_gemm4096 python tools/gemm_debug.py perf > $OUT/${TAG}_gemm4096.log 2>&1
export_rep ${TAG}_gemm4096 gemm_tc
ls -la $OUT
du -sh $OUT
