#!/bin/bash
# r02 trip 12 (8 GPUs): peer-push exchange: gemm e2e (+trace) and csrmm cfg-3 shared-B at N=8
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t12; mkdir -p $OUT
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --no-cpu --extra pcie,cfg3 --steps 2 > $OUT/trace_8gpu.txt 2>&1
tail -1 $OUT/trace_8gpu.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e'].get('spot_rel_err'))
print(json.dumps(d['extra']['pcie']))
v = d['extra']['csrmm_cfg3']
for k in ('e2e', 'e2e_shared_b', 'parity', 'error', 'trace'): print(k, json.dumps(v.get(k))[:400])"
