#!/bin/bash
# compute-sanitizer memcheck + racecheck over every kernel family at small sizes.
# Usage (GPU box): bash tools/sanitize_gpu.sh [out_dir]   -- logs land in gpurun_out/sanitize/
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
for job in ${JOBS:-small uvd_tma uvd_direct kron_ts kron_ss kron_pair kron_stream splu vec}; do
  for tool in memcheck racecheck; do
    timeout 420 $SAN --tool $tool --print-limit 20 --log-file "$OUT/${tool}_${job}.log" \
      python tools/sanitize_cases.py $job > "$OUT/${tool}_${job}.out" 2>&1
    echo "$tool $job rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${tool}_${job}.log" | tail -1)" | tee -a "$OUT/summary.txt"
  done
done
