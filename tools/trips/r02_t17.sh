#!/bin/bash
# r02 trip 17 (4 GPUs): the one-process multi-GPU path at 4 devices (a shard without rows used to skip the exchange
# hand-shake and hang its peers), uneven shards under torchrun, BOF_GPUS drivers, out-of-core k-means
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t17; mkdir -p $OUT
timeout 150 python tools/mgpu_check.py > $OUT/mgpu_check.txt 2>&1; echo "mgpu_check rc=$?" | tee -a $OUT/mgpu_check.txt; tail -12 $OUT/mgpu_check.txt
timeout 330 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kmeans.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -25 > $OUT/tests.txt; tail -25 $OUT/tests.txt
