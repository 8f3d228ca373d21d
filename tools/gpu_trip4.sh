#!/bin/bash
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary4.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary4.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary4.txt; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-700 | tee -a $OUT/summary4.txt; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
TAILN=30 run suite python tools/bench_suite.py --only cfg4,cfg5,cfg1 --out $OUT/suite.json
run bench python bench.py --steps 3 --warmup 3
# full-size gemm kernel: one ncu --set full capture for the roofline 'traffic' field
run ncu_gemm_full ncu --set full --clock-control none -k regex:gemm3xtf32 -s 1 -c 1 -f -o $OUT/prof_gemm32k python bench.py --steps 1 --warmup 1 --no-cpu --no-extra
ncu -i $OUT/prof_gemm32k.ncu-rep --page raw --csv > $OUT/prof_gemm32k_raw.csv 2>/dev/null
rm -f $OUT/prof_gemm32k.ncu-rep
run ncu_radix ncu --set full --clock-control none --import-source on -k "regex:radix_scatter|radix_hist|kmeans_partial|kmeans_combine" -c 6 -f -o $OUT/prof_radix2 python tools/prof_targets.py csrcsc,kmeans
ncu -i $OUT/prof_radix2.ncu-rep --page raw --csv > $OUT/prof_radix2_raw.csv 2>/dev/null
ncu -i $OUT/prof_radix2.ncu-rep --page source --csv -k regex:radix_scatter > $OUT/prof_radix2_scatter_source.csv 2>/dev/null
rm -f $OUT/prof_radix2.ncu-rep
du -sm $OUT | tee -a $OUT/summary4.txt
