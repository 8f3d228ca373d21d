#!/bin/bash
# r02 trip 16 (4 GPUs): the full bench line at N=4 (the one world size not run yet) + multi-GPU tests incl. BOF_GPUS drivers
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t16; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sparse.py tests/test_gpu_csrcsc.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -12 > $OUT/tests.txt; tail -4 $OUT/tests.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 ) > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
tail -4 $OUT/bench_4gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_t16/bench_4gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("x_of_duplex_bound"), d["e2e"].get("x_of_bound"))
        print(json.dumps(d["extra"].get("pcie")))
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "bench_seconds", "error", "trace")})[:300])
            for sub in ("e2e", "e2e_shared_b"):
                if sub in v: print("   ", sub, json.dumps(v[sub])[:420])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
