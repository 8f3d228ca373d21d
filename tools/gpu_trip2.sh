#!/bin/bash
# Trip 2: per-config suite, ncu launch list of the bench command, ncu --set full of the hot kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary2.txt; timeout ${TMO:-1500} "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary2.txt; tail -n ${TAILN:-12} $OUT/$name.log | cut -c1-600 | tee -a $OUT/summary2.txt; }
: > $OUT/summary2.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_idle.csv 2>&1
TAILN=40 run suite python tools/bench_suite.py --out $OUT/suite.json
run launches_bench ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu
run ncu_full ncu --set full --clock-control none --import-source on -k "regex:gemm3xtf32|spmm_csr|radix_scatter|radix_hist|spmv_csr|kmeans_segment|split_planes_kmajor|scan_apply" -c 40 -f -o $OUT/prof_r01 python tools/prof_targets.py
ls -la $OUT | tee -a $OUT/summary2.txt
