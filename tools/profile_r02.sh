#!/usr/bin/env bash
# ncu evidence for round 2 (run under gpurun on ONE B200; see B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh r02'
# Launch lists of both headline workloads (eager: --no-graphs, so that every kernel is a launch ncu can name) and
# --set full captures of the UVd sweeps + mid kernel, of ALL 36 tcgen05 GEMM launches of one Kron step (6 layers per grouped
# launch: same kernels and tile shapes, keeps ncu's save/restore small) and of one dense 4096^3 product.  Only CSV/text is
# left in gpurun_out/ (capped at 64 MiB); tools/summarise_profiles.py <tag> writes profiles/<tag>_*.
set -u
TAG=${1:-r02}
OUT=gpurun_out
TMP=/tmp/psgd_prof
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
UVD="python bench.py --workload uvd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-separate --no-parity --no-graphs"
KRON="python bench.py --workload kron --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
OURS='gram_sweep|map_sweep|uvd_mid|reduce_partials|uvd_small|maxabs2|zero_small|exchange|gemm_tc|gemm_simt|tri_inv|tri_scan|zero_flags|balance|rescale|trsm'
export_rep() {   # $1 = rep basename, $2 = kernel regex for the source page
  ncu -i $TMP/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $TMP/$1.ncu-rep --page source --csv --kernel-name regex:"$2" --launch-count 1 > $OUT/$1_source.csv 2>/dev/null
  rm -f $TMP/$1.ncu-rep
}
# 1. launch lists: 3 warm-up steps skipped (UVd: 5 launches per step; Kron: 43 per step incl. the apply call)
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 15 -c 10 --csv --log-file $OUT/${TAG}_uvd_launches.csv $UVD > $OUT/${TAG}_uvd_launches.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 129 -c 43 --csv --log-file $OUT/${TAG}_kron_launches.csv $KRON > $OUT/${TAG}_kron_launches.log 2>&1
# 2. full captures
$NCU --set full --import-source on -k regex:'gram_sweep|map_sweep|uvd_mid' -s 15 -c 5 -f -o $TMP/${TAG}_uvd_full $UVD > $OUT/${TAG}_uvd_full.log 2>&1
export_rep ${TAG}_uvd_full gram_sweep
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 108 -c 36 -f -o $TMP/${TAG}_kron_full $KRON --layers 6 > $OUT/${TAG}_kron_full.log 2>&1
export_rep ${TAG}_kron_full gemm_tc
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $TMP/${TAG}_gemm4096 python tools/gemm_debug.py perf > $OUT/${TAG}_gemm4096.log 2>&1
export_rep ${TAG}_gemm4096 gemm_tc
ls -la $OUT | grep ${TAG}_
du -sh $OUT
