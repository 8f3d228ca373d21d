#!/bin/bash
# r02 trip 6 (2 GPUs): gemm e2e trace with slab-download marks
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t06; mkdir -p $OUT
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --no-extra --no-cpu --steps 2 > $OUT/trace_2gpu.txt 2>&1
grep -c "" $OUT/trace_2gpu.txt
tail -1 $OUT/trace_2gpu.txt | cut -c1-300
