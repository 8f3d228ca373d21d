#!/bin/bash
# N-GPU trip (N = $1) on the final code: bench.py and cfg-3 csrmm in both B-distribution modes.
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary22_${N}gpu.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; timeout 420 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $S; grep "^{" $OUT/$name.log | cut -c1-2500 | tee -a $S; tail -n 3 $OUT/$name.log | grep -v "^{" | cut -c1-300 | tee -a $S; }
nvidia-smi -L | tee -a $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
run csrmm_${N}gpu $TR tools/csrmm_multi.py
run bench_${N}gpu $TR bench.py --gpus $N --steps 3 --warmup 3 --no-extra
