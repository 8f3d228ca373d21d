#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary7.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary7.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary7.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-900 | tee -a $OUT/summary7.txt; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
TAILN=20 run suite python tools/bench_suite.py --only pcie,cfg4,cfg3 --out $OUT/suite.json
run bench python bench.py --steps 3 --warmup 3
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
