#!/bin/bash
# Trip 23: resident-CSR handle (tests + bench), csrmm_pmem driver, bench.py with the NVML clock sampler.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary23.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; timeout ${TMO:-900} "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $S; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-3000 | tee -a $S; }
run tests_resident python -m pytest tests/test_gpu_resident.py tests/test_gpu_drivers.py -m gpu -q --tb=short -p no:cacheprovider -x
run resident_bench python tools/resident_bench.py
run bench python bench.py --steps 3 --warmup 3 --no-cpu --no-extra
