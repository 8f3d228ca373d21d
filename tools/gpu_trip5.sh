#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary5.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary5.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary5.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-700 | tee -a $OUT/summary5.txt; }
run tests_gemm python -m pytest tests/test_gpu_gemm.py tests/test_gpu_drivers.py -m gpu -q --tb=short -p no:cacheprovider -x
TAILN=12 run sync_sweep python tools/bench_suite.py --only sync --out $OUT/suite_sync.json
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
run ncu_sync   ncu --metrics $M --clock-control none -k regex:gemm3xtf32 -s 1 -c 1 --csv --log-file $OUT/ncu_gemm32k_sync.csv python tools/prof_targets.py gemm32k
run ncu_nosync ncu --metrics $M --clock-control none -k regex:gemm3xtf32 -s 1 -c 1 --csv --log-file $OUT/ncu_gemm32k_nosync.csv python tools/prof_targets.py gemm32k_nosync
run bench python bench.py --steps 3 --warmup 3
du -sm $OUT | tee -a $OUT/summary5.txt
