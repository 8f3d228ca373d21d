#!/bin/bash
# r02 trip 18c (last ~2 GPU-minutes of the round): the rest of tests/test_gpu_sparse.py after the fixed test, then a short bench line
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t18; mkdir -p $OUT
timeout 75 python -m pytest tests/test_gpu_sparse.py -m gpu -q -p no:cacheprovider > $OUT/tests_sparse.txt 2>&1; echo "pytest rc=$?" >> $OUT/tests_sparse.txt
tail -5 $OUT/tests_sparse.txt
timeout 45 python bench.py --no-cpu --extra cfg4 --steps 2 > $OUT/bench_short.json 2> $OUT/bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_t18/bench_short.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "error")})[:200])
PY
