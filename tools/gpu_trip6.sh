#!/bin/bash
# 2-GPU trip: tests of the host gemm change, then strong-scaling bench and sharded k-means under torchrun.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary6.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary6.txt; timeout 1200 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary6.txt; tail -n ${TAILN:-6} $OUT/$name.log | cut -c1-900 | tee -a $OUT/summary6.txt; }
nvidia-smi -L | tee -a $OUT/summary6.txt
run tests_gemm python -m pytest tests/test_gpu_gemm.py tests/test_gpu_drivers.py -m gpu -q --tb=short -p no:cacheprovider -x
run bench_1gpu python bench.py --steps 3 --warmup 3 --no-extra
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run bench_ref_2gpu $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1
run bench_2gpu $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-extra
run kmeans_2gpu $TR tools/kmeans_multi.py --check
run kmeans_1gpu python tools/kmeans_multi.py --check
