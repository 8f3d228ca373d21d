#!/bin/bash
# r02 trip 7: SpMM row splitting (parity on power-law matrices, timing with / without), cfg-1 / cfg-3 regression check
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t07; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_resident.py tests/test_gpu_ref_parity.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > $OUT/tests.txt
tail -4 $OUT/tests.txt
timeout 600 python tools/bench_spmm_skew.py > $OUT/spmm_skew.txt 2>&1
timeout 600 python tools/bench_spmm_skew.py --k 256 --nnz 33554432 >> $OUT/spmm_skew.txt 2>&1
cat $OUT/spmm_skew.txt
timeout 900 python bench.py --no-cpu --extra cfg1,cfg3 --steps 2 > $OUT/bench_cfg13.json 2> $OUT/bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_t07/bench_cfg13.json"):
    if l.startswith("{"):
        d = json.loads(l)
        for k, v in d["extra"].items():
            print(k, v.get("value"), v.get("ms"), v.get("error"), v.get("parity"))
PY
