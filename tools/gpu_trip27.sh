#!/bin/bash
# Trip 27 (2 GPUs): the driver's round-end sequence -- both arms at N=1 and N=2, default flags.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary27.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; t0=$SECONDS; timeout 900 "$@" > $OUT/$name.log 2>&1; echo "exit $? wall $((SECONDS - t0)) s" | tee -a $S; grep "^{" $OUT/$name.log | cut -c1-2600 | tee -a $S; tail -n 2 $OUT/$name.log | grep -v "^{" | cut -c1-300 | tee -a $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
run ref_1gpu python bench.py --impl reference --gpus 1 --steps 3 --warmup 3
run ours_1gpu python bench.py
run ref_2gpu $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3
run ours_2gpu $TR bench.py --gpus 2 --steps 3 --warmup 3
