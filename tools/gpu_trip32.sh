#!/bin/bash
# Trip 32: async device->host staging (drainer jobs): full GPU suite + file-backed driver bench.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary32.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; t0=$SECONDS; timeout ${TMO:-1200} "$@" > $OUT/$name.log 2>&1; echo "exit $? wall $((SECONDS - t0)) s" | tee -a $S; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-1500 | tee -a $S; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run drivers_big python tools/driver_bench.py --rows 8388608 --gemm 32768
