#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary12.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary12.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary12.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-1100 | tee -a $OUT/summary12.txt; }
g++ -O2 -std=c++17 -pthread tools/pooltest.cpp -o /tmp/pooltest && /tmp/pooltest 2>&1 | tee -a $OUT/summary12.txt
BOF_AB=1 run drivers python tools/driver_bench.py
run drivers_big python tools/driver_bench.py --rows 8388608 --gemm 32768
