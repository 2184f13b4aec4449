#!/usr/bin/env bash
# ncu evidence for the streaming (HBM-bound) paths, cheap enough to re-run per kernel change (~1 min of GPU time):
#   gpurun --timeout 900 -- 'bash tools/profile_uvd.sh r01b'
# UVd headline workload (fused update+apply form): launch list of the timed steps + one --set full capture of the three
# sweeps; SPLU passes and the (normalization, scaling) reducing kernels: one --set full capture each.  Everything is
# exported to CSV on the box (the .ncu-rep files are deleted: gpurun_out/ is capped at 64 MiB);
# tools/summarise_profiles.py <tag> then writes profiles/<tag>_*.
set -u
TAG=${1:-r01b}
OUT=gpurun_out
TMP=/tmp/psgd_prof
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
UVD="python bench.py --workload uvd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-separate"
OURS='gram_sweep|map_sweep|reduce_partials|uvd_small|d_update|maxabs2|balance|zero_small|exchange'
export_rep() {   # $1 = rep basename, $2 = kernel regex for the source page, $3 = suffix of the source file
  ncu -i $TMP/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $TMP/$1.ncu-rep --page source --csv --kernel-name regex:"$2" --launch-count 1 > $OUT/${1%_full}_${3}source.csv 2>/dev/null
}
# launch list: the 2 timed steps (3 warm-up steps x 7 launches are skipped)
$NCU --metrics gpu__time_duration.sum -k regex:"$OURS" -s 21 --csv --log-file $OUT/${TAG}_uvd_launches.csv $UVD > $OUT/${TAG}_uvd_launches.log 2>&1
# full capture of the three sweeps of one step (+ per-instruction source pages of the Gram and the fused map sweep)
$NCU --set full --import-source on -k regex:'gram_sweep|map_sweep' -s 9 -c 3 -f -o $TMP/${TAG}_uvd_full $UVD > $OUT/${TAG}_uvd_full.log 2>&1
export_rep ${TAG}_uvd_full gram_sweep full_
ncu -i $TMP/${TAG}_uvd_full.ncu-rep --page source --csv --kernel-name regex:map_sweep --launch-count 1 > $OUT/${TAG}_uvd_map_source.csv 2>/dev/null
rm -f $TMP/${TAG}_uvd_full.ncu-rep
# SPLU: the four passes of the second update at n = 5e7, r = 10
$NCU --set full --import-source on -k regex:'pass._kernel' -s 7 -c 4 -f -o $TMP/${TAG}_splu_full python tools/splu_probe.py 5e7 2 > $OUT/${TAG}_splu_full.log 2>&1
export_rep ${TAG}_splu_full pass4_kernel full_
rm -f $TMP/${TAG}_splu_full.ncu-rep
# (normalization, scaling) update at [8192, 8192]: the two reducing kernels
$NCU --set full --import-source on -k regex:'ns_stats_kernel|col_wsum_kernel' -s 2 -c 2 -f -o $TMP/${TAG}_ns_full python tools/ns_probe.py > $OUT/${TAG}_ns_full.log 2>&1
export_rep ${TAG}_ns_full ns_stats_kernel full_
rm -f $TMP/${TAG}_ns_full.ncu-rep
ls -la $OUT | grep ${TAG}_
