#!/bin/bash
# r02 trip 18 (1 GPU, last of the round's budget): full GPU suite after the out-of-core k-means, column-panel csrmm,
# collective-call and radix changes; then the bench line without the file-backed extra (53 s) to stay inside the budget
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t18; mkdir -p $OUT
STEP=${1:-all}
if [ "$STEP" = tests ] || [ "$STEP" = all ]; then
timeout 170 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > $OUT/tests.txt; tail -6 $OUT/tests.txt
fi
if [ "$STEP" = tests ]; then exit 0; fi
timeout 100 python bench.py --extra pcie,cfg1,cfg4,cfg5 > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_t18/bench_1gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "error")})[:200])
PY
tail -3 $OUT/bench_1gpu.err | cut -c1-300
