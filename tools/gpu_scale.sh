#!/usr/bin/env bash
# bench.py under torchrun on N GPUs of one box (what the driver's scaling run does): tools/gpu_scale.sh N
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 3 > $OUT/scale_$N.json 2> $OUT/scale_$N.err; echo "rc=$?"
python - "$OUT/scale_$N.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['cross_gpu_exchange'], d['config']['cuda_graphs'], d['gpu_launches'], d['e2e']['value'])
print('kron', d['kron']['value'], d['kron']['ms_per_step'], d['kron']['e2e']['value'])
PY
tail -3 $OUT/scale_$N.err
