#!/usr/bin/env bash
# Kron stack step time for each triangular-solve base-block size (tuning aid).
for b in 128 256 512 1024; do
  PSGD_TRSM_BASE=$b python bench.py --workload kron --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > /tmp/kb.json
  python -c "
import json; d=json.load(open('/tmp/kb.json')); print('trsm_base', $b, d['value'], 'steps/s', d['ms_per_step'], 'ms', d['gpu_launches'], 'launches')"
done
