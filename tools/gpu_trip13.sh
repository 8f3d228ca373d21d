#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary13.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary13.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary13.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-1100 | tee -a $OUT/summary13.txt; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run drivers python tools/driver_bench.py
run drivers_big python tools/driver_bench.py --rows 8388608 --gemm 32768
run bench python bench.py --steps 3 --warmup 3
