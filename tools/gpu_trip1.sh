#!/bin/bash
# First GPU bring-up: every test group in its own process (a trap in one kernel poisons its CUDA context only).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary.txt; timeout 900 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary.txt; tail -n 15 $OUT/$name.log | tee -a $OUT/summary.txt; }
: > $OUT/summary.txt
run sparse  python -m pytest tests/test_gpu_sparse.py -m gpu -q --tb=short -p no:cacheprovider
run csrcsc  python -m pytest tests/test_gpu_csrcsc.py -m gpu -q --tb=short -p no:cacheprovider
run gemm_ffma python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -p no:cacheprovider -k "ffma"
run gemm_tc1 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -p no:cacheprovider -k "tc1" -s
run gemm_tc2 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -p no:cacheprovider -k "tc2" -s
run gemm_rest python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -p no:cacheprovider -k "not tc1 and not tc2 and not ffma" -s
run kmeans  python -m pytest tests/test_gpu_kmeans.py -m gpu -q --tb=short -p no:cacheprovider
run smoke   python -c "import __graft_entry__ as g; g.smoke()"
run bench_small python bench.py --size 8192 --steps 3 --warmup 3 --no-cpu
run bench_full  python bench.py --steps 3 --warmup 3
