#!/bin/bash
# r02 trip 3: profile the new radix kernels (launch list restricted to them + one --set full capture)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t03; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"radix_|scan_|super_table|offsets_" -c 40 --csv --log-file $OUT/csrcsc_launches.csv \
  python tools/bench_csrcsc.py --rows 2097152 --bits 12,8 --iters 1 > $OUT/ncu_stdout.txt 2>&1
python tools/launch_list.py $OUT/csrcsc_launches.csv --per-launch kernel > $OUT/csrcsc_per_launch.txt 2>&1
cat $OUT/csrcsc_per_launch.txt | head -60
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"radix_" -c 4 -o $OUT/radix_full \
  python tools/bench_csrcsc.py --rows 2097152 --bits 12 --iters 1 > $OUT/ncu_full_stdout.txt 2>&1
ls -la $OUT
