#!/usr/bin/env bash
# ncu evidence for the UVd headline workload only (fused update+apply form), cheap enough to re-run per kernel change:
#   gpurun --timeout 900 -- 'bash tools/profile_uvd.sh r01b'
# Same outputs as the UVd part of tools/profile_gpu.sh: launch list + one --set full capture of the sweeps, exported to
# CSV on the box (the .ncu-rep is deleted: gpurun_out/ is capped at 64 MiB).  tools/summarise_profiles.py <tag> then
# writes profiles/<tag>_*.
set -u
TAG=${1:-r01b}
OUT=gpurun_out
TMP=/tmp/psgd_prof
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
UVD="python bench.py --workload uvd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-separate"
OURS='gram_sweep|map_sweep|reduce_partials|uvd_small|d_update|maxabs2|balance|zero_small|exchange'
# launch list: the 2 timed steps (3 warm-up steps x 7 launches are skipped)
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 21 --csv --log-file $OUT/${TAG}_uvd_launches.csv $UVD > $OUT/${TAG}_uvd_launches.log 2>&1
# full capture of the three sweeps of one step (+ per-instruction source page of the dominant one)
$NCU --set full --import-source on -k regex:'gram_sweep|map_sweep' -s 9 -c 3 -f -o $TMP/${TAG}_uvd_full $UVD > $OUT/${TAG}_uvd_full.log 2>&1
ncu -i $TMP/${TAG}_uvd_full.ncu-rep --page raw --csv > $OUT/${TAG}_uvd_full_raw.csv 2>/dev/null
ncu -i $TMP/${TAG}_uvd_full.ncu-rep --page source --csv --kernel-name regex:'map_sweep_kernel<10, (9|10)>' --launch-count 1 > $OUT/${TAG}_uvd_full_source.csv 2>/dev/null
rm -f $TMP/${TAG}_uvd_full.ncu-rep
ls -la $OUT | head -30
