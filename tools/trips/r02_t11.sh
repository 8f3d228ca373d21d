#!/bin/bash
# r02 trip 11 (2 GPUs): peer-push exchange (copy engines + flags): parity (one process and torchrun), N=2 bench e2e + trace
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t11; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > $OUT/tests.txt; tail -5 $OUT/tests.txt
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --no-extra --no-cpu --steps 2 > $OUT/trace_2gpu.txt 2>&1
tail -1 $OUT/trace_2gpu.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e'].get('spot_rel_err'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --no-cpu --extra cfg3 --steps 2 2>$OUT/bench_cfg3.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); v = d['extra']['csrmm_cfg3']
for k in ('e2e', 'e2e_shared_b', 'parity', 'error', 'trace'): print(k, json.dumps(v.get(k))[:400])"
