#!/bin/bash
# Trip 37: A/B of MADV_POPULATE on the pageable (file-backed) staging path.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
for mode in 1 0; do
  echo "=== BOF_POPULATE=$mode" | tee -a $OUT/populate_ab.txt
  BOF_POPULATE=$mode timeout 400 python tools/driver_bench.py --rows 8388608 --gemm 32768 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['driver_says'])" | tee -a $OUT/populate_ab.txt
done
timeout 200 python -m pytest tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2 | tee -a $OUT/populate_ab.txt
