#!/bin/bash
# Trip 3: ncu launch list of the bench command + ncu --set full per kernel group, exported to CSV on the box
# (gpurun_out is capped at 64 MiB, so reports are kept small and summarised as text).
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary3.txt
run() { name=$1; shift; echo "=== $name" | tee -a $OUT/summary3.txt; timeout 1500 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/summary3.txt; tail -n 6 $OUT/$name.log | cut -c1-300 | tee -a $OUT/summary3.txt; }
run launches_bench ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu
prof() { tag=$1; regex=$2; count=$3; targets=$4
  run ncu_$tag ncu --set full --clock-control none --import-source on -k "regex:$regex" -c $count -f -o $OUT/prof_$tag python tools/prof_targets.py $targets
  ncu -i $OUT/prof_$tag.ncu-rep --page raw --csv > $OUT/prof_${tag}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$tag.ncu-rep --page details --csv > $OUT/prof_${tag}_details.csv 2>/dev/null
}
prof gemm "gemm3xtf32" 2 gemm
prof spmm "spmm_csr|spmv_csr" 8 spmm,spmv
prof radix "radix_scatter|radix_hist|scan_apply|kmeans_segment|expand_rows|segment_offsets" 8 csrcsc
prof kmeans "gemm3xtf32|kmeans_segment|radix_scatter" 6 kmeans
# source-level hot spots for the two kernels that matter most
ncu -i $OUT/prof_radix.ncu-rep --page source --csv -k regex:radix_scatter > $OUT/prof_radix_scatter_source.csv 2>/dev/null
ncu -i $OUT/prof_gemm.ncu-rep --page source --csv > $OUT/prof_gemm_source.csv 2>/dev/null
du -sh $OUT/* | tee -a $OUT/summary3.txt
# keep the merged directory under the cap: drop the largest reports if needed
total=$(du -sm $OUT | cut -f1); if [ "$total" -gt 55 ]; then rm -f $OUT/prof_radix.ncu-rep $OUT/prof_kmeans.ncu-rep; fi
total=$(du -sm $OUT | cut -f1); if [ "$total" -gt 55 ]; then rm -f $OUT/prof_spmm.ncu-rep; fi
du -sm $OUT | tee -a $OUT/summary3.txt
