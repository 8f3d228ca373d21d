#!/bin/bash
# r02 trip 14: full GPU suite; ncu traffic per launch for every config at full size; --set full captures of the top kernels
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t14; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=8 2>&1 | tail -25 > $OUT/tests.txt; tail -14 $OUT/tests.txt
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
  --csv --log-file $OUT/launches_all.csv python tools/prof_r02.py cfg1,skew,cfg3,cfg4,cfg5,gemm > $OUT/prof_stdout.txt 2>&1
python tools/launch_list.py $OUT/launches_all.csv --per-launch kernel > $OUT/per_launch_all.txt 2>&1
cat $OUT/per_launch_all.txt | head -80
for spec in "cfg1:spmm_csr_rm_vec" "cfg3:spmm_csr_rm_vec" "skew:spmm_long_rows" "cfg4:radix_scatter" "cfg5:gemm3xtf32" "gemm:gemm3xtf32"; do
  cfg=${spec%%:*}; k=${spec##*:}
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k -c 1 -o $OUT/full_${cfg}_${k} \
    python tools/prof_r02.py $cfg > $OUT/full_${cfg}_stdout.txt 2>&1
done
ls -la $OUT
