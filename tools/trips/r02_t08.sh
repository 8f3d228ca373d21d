#!/bin/bash
# r02 trip 8 (2 GPUs): gemm e2e trace with slab-download marks + quick sparse regression (split counters self-reset)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t08; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3 > $OUT/tests.txt; cat $OUT/tests.txt
timeout 300 python tools/bench_spmm_skew.py > $OUT/spmm_skew.txt 2>&1; cat $OUT/spmm_skew.txt
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --no-extra --no-cpu --steps 2 > $OUT/trace_2gpu.txt 2>&1
grep -c "" $OUT/trace_2gpu.txt
tail -1 $OUT/trace_2gpu.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'])"
