#!/usr/bin/env bash
# 2-GPU validation: sharded CUDA paths + the bench under torchrun
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/c11_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/c11_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 > $OUT/c11_bench2.json 2> $OUT/c11_bench2.err; echo "bench2 rc=$?" >> $OUT/c11_status.txt
cat $OUT/c11_status.txt; tail -5 $OUT/c11_pytest.log; tail -3 $OUT/c11_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c11_bench2.json')); print(d['value'], d['ms_per_step'], d['config']['cross_gpu_exchange'], d['config']['cuda_graphs'], d['gpu_launches'], [(k['kernel'],k['avg_ms']) for k in d['kernels']], d['e2e'])
print('kron', d['kron']['value'], d['kron']['ms_per_step'])
PY
