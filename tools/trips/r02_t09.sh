#!/bin/bash
# r02 trip 9 (2 GPUs): gemm e2e trace after limiting NCCL CTAs / reserving SMs; dist parity
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t09; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3 > $OUT/tests.txt; cat $OUT/tests.txt
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --no-extra --no-cpu --steps 2 > $OUT/trace_2gpu.txt 2>&1
tail -1 $OUT/trace_2gpu.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'])"
for c in 2 8; do
BOF_NCCL_MAX_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --no-extra --no-cpu --steps 2 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('max_ctas $c: value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'])"
done
