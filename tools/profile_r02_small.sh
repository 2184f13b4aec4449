#!/bin/bash
# Round-2 evidence for the small-layer engine and the dense preconditioner (GPU box):
#   CUPTI per-kernel tables, clock64 phase stamps, one ncu --set full capture per new kernel.
# Usage: bash tools/profile_r02_small.sh [out_dir]
OUT=${1:-gpurun_out/r02_small}
mkdir -p "$OUT"
python tools/small_kernels.py 257 120 2>&1 | grep -v -i warn > "$OUT/small_kernels_257x120.txt"
python tools/small_time.py 2>&1 | grep -v -i warn > "$OUT/small_time_cfg1.txt"
python tools/trsm_stamps.py 257 120 > "$OUT/trsm_stamps_257x120.txt" 2>&1
python tools/trsv_stamps.py 512 > "$OUT/trsv_stamps_512.txt" 2>&1
python tools/dense_kernels.py 8192 2>&1 | grep -v -i warn > "$OUT/dense_kernels_8192.txt"
python tools/dense_kernels.py 1000 2>&1 | grep -v -i warn > "$OUT/dense_kernels_1000.txt"
python tools/nmt_kernels.py 2>&1 | grep -v -i warn > "$OUT/nmt_kernels.txt"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:scan_kernel -s 4 -c 2 -o "$OUT/dense_scan" python tools/dense_kernels.py 8192 > "$OUT/ncu_dense.log" 2>&1
timeout 300 $NCU -k regex:"trsm_panel_kernel|gemm_small_kernel" -s 24 -c 6 -o "$OUT/small_layer" python tools/small_kernels.py 257 120 > "$OUT/ncu_small.log" 2>&1
for f in dense_scan small_layer; do
  ncu -i "$OUT/$f.ncu-rep" --page raw --csv > "$OUT/${f}_raw.csv" 2>/dev/null
done
ls -la "$OUT"
