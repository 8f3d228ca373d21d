#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary11.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary11.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary11.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-900 | tee -a $OUT/summary11.txt; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run drivers python tools/driver_bench.py
run drivers_big python tools/driver_bench.py --rows 8388608 --gemm 32768
