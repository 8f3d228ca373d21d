#!/bin/bash
# r02 trip 2: new wide-digit csrcsc: parity tests, k-means (reduce uses the same sort), cfg-4 timing + launch list
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t02; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_csrcsc.py tests/test_gpu_kmeans.py tests/test_gpu_ref_parity.py tests/test_gpu_resident.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > $OUT/tests.txt
tail -3 $OUT/tests.txt
timeout 600 python tools/bench_csrcsc.py --bits 12,8 > $OUT/csrcsc_bench.txt 2>&1
cat $OUT/csrcsc_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $OUT/csrcsc_launches.csv python tools/bench_csrcsc.py --bits 12 --iters 1 > $OUT/ncu_stdout.txt 2>&1
python tools/launch_list.py $OUT/csrcsc_launches.csv 2>/dev/null | tail -30
