#!/bin/bash
# r02 trip 10 (8 GPUs): full bench line at N=8 (all configs), then a gemm-only trace
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t10; mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv | wc -l > $OUT/box.txt; nproc >> $OUT/box.txt; free -g | head -2 >> $OUT/box.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 ) > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
tail -4 $OUT/bench_8gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_t10/bench_8gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"])
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "bench_seconds", "error", "trace", "h2d_gbs_per_gpu", "h2d_gbs_aggregate")}))
            for sub in ("e2e", "e2e_shared_b", "parity"):
                if sub in v: print("   ", sub, json.dumps(v[sub])[:420])
PY
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --no-extra --no-cpu --steps 2 > $OUT/trace_8gpu.txt 2>&1
tail -1 $OUT/trace_8gpu.txt | cut -c1-200
