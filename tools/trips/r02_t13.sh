#!/bin/bash
# r02 trip 13: page-locked file mappings (drivers on /dev/shm files) A/B, deterministic csrgemv 'T' tests + cfg-4 timing
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t13; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_drivers.py tests/test_gpu_resident.py tests/test_gpu_staging.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > $OUT/tests.txt; tail -4 $OUT/tests.txt
for pin in 1 0; do
  echo "=== BOF_PIN_MAPPINGS=$pin" | tee -a $OUT/driver_bench.txt
  BOF_PIN_MAPPINGS=$pin BOF_TRACE=${pin} timeout 900 python tools/driver_bench.py --rows 8388608 --gemm 32768 2>$OUT/driver_err_$pin.txt | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k not in ('config',)}, d['config'][:40])" | tee -a $OUT/driver_bench.txt
done
grep "page-locked\|cudaHostRegister" $OUT/driver_err_1.txt | head -20 | tee -a $OUT/driver_bench.txt
timeout 900 python bench.py --no-cpu --extra cfg4 --steps 2 2> $OUT/bench.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d['extra'].items():
    print(k, json.dumps({kk: v.get(kk) for kk in ('value', 'unit', 'ms', 'error', 'atomic_variant', 'parity')})[:600])"
