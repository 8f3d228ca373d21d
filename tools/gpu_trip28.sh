#!/bin/bash
# Trip 28 (N GPUs): A/B of the NUMA binding of the ranks' pinned host buffers on the e2e leg.
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary28_${N}gpu.txt
: > $S
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)") 2>&1 | tee -a $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; t0=$SECONDS; timeout 900 "$@" > $OUT/$name.log 2>&1; echo "exit $? wall $((SECONDS - t0)) s" | tee -a $S; grep "^{" $OUT/$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(json.dumps({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'e2e', 'clocks', 'host')}))" | tee -a $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556"
run bind_${N}gpu $TR bench.py --gpus $N --steps 3 --warmup 3 --no-extra
run nobind_${N}gpu $TR bench.py --gpus $N --steps 3 --warmup 3 --no-extra --no-numa-bind
run bind_1gpu python bench.py --steps 3 --warmup 3 --no-extra --no-cpu
run nobind_1gpu python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-numa-bind
