#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary9.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary9.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary9.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-700 | tee -a $OUT/summary9.txt; }
TAILN=30 run tests_gemm python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kmeans.py tests/test_gpu_csrcsc.py -m gpu -q --tb=short -p no:cacheprovider -s
TAILN=10 run split_sweep python tools/bench_suite.py --only split --out $OUT/suite_split.json
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
BOF_SPLIT=2 run ncu_hyb ncu --metrics $M --clock-control none -k regex:gemm3xtf32 -s 1 -c 1 --csv --log-file $OUT/ncu_gemm32k_hybrid.csv python tools/prof_targets.py gemm32k_hyb
