#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary10.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary10.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary10.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-900 | tee -a $OUT/summary10.txt; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run bench python bench.py --steps 3 --warmup 3
run drivers python tools/driver_bench.py
TAILN=12 run suite python tools/bench_suite.py --only cfg5,cfg4 --out $OUT/suite.json
df -h /dev/shm | tee -a $OUT/summary10.txt
