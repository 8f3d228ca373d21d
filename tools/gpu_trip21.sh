#!/bin/bash
# Trip 21: final-code evidence refresh on one GPU: full GPU tests, smoke, SpMM variant sweep (incl. the TMA-staged
# kernel), bench, ncu launch list of the bench command, ncu --set full of every hot kernel (CSV exports only).
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary21.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; timeout ${TMO:-1200} "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $S; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-1600 | tee -a $S; }
run tests_sparse python -m pytest tests/test_gpu_sparse.py -m gpu -q --tb=short -p no:cacheprovider -x
TAILN=10 run spmm_variants python tools/spmm_variants.py
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')"
run bench python bench.py --steps 3 --warmup 3
grep "^{" $OUT/bench.log > $OUT/bench.json
run launches_bench ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu
prof() { tag=$1; regex=$2; count=$3; targets=$4
  TMO=900 run ncu_$tag ncu --set full --clock-control none --import-source on -k "regex:$regex" -c $count -f -o $OUT/prof_$tag python tools/prof_targets.py $targets
  ncu -i $OUT/prof_$tag.ncu-rep --page raw --csv > $OUT/prof_${tag}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$tag.ncu-rep --page details --csv > $OUT/prof_${tag}_details.csv 2>/dev/null
  rm -f $OUT/prof_$tag.ncu-rep
}
prof gemm32k_hyb "gemm3xtf32" 1 gemm32k_hyb
prof spmm "spmm_csr|spmv_csr" 8 spmm,spmv
BOF_SPMM_VARIANT=6 prof spmm_tma "spmm_csr" 4 spmm
prof radix "radix_scatter|radix_hist|scan_|expand_rows|segment_offsets" 10 csrcsc
prof kmeans "gemm3xtf32|kmeans_|radix_scatter|split_planes" 12 kmeans
for f in $OUT/prof_*_raw.csv; do python tools/ncu_summary.py $f >> $OUT/ncu_full_summary21.txt 2>&1; done
du -sm $OUT | tee -a $S
