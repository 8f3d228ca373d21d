#!/bin/bash
# r02 trip 5 (2 GPUs): multi-GPU parity (one process / torchrun), then bench.py at N=2 and N=1 gemm-only e2e
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t05; mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv > $OUT/gpus.txt; nproc >> $OUT/gpus.txt; free -g | head -2 >> $OUT/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gemm.py tests/test_gpu_kmeans.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > $OUT/tests.txt
tail -5 $OUT/tests.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 ) > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -4 $OUT/bench_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_t05/bench_2gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"])
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "bench_seconds", "error", "trace")}))
            for sub in ("e2e", "e2e_shared_b", "parity"):
                if sub in v: print("   ", sub, json.dumps(v[sub])[:500])
PY
BOF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --no-extra --no-cpu --steps 2 > $OUT/trace_2gpu.txt 2>&1
grep -c "" $OUT/trace_2gpu.txt
