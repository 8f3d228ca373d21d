#!/bin/bash
# r02 trip 15: supertile radix (tests, cfg-4 timing, per-launch), then the full default bench line + reference arm
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t15; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_csrcsc.py tests/test_gpu_kmeans.py tests/test_gpu_sparse.py tests/test_gpu_ref_parity.py tests/test_gpu_resident.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -12 > $OUT/tests.txt; tail -3 $OUT/tests.txt
timeout 600 python tools/bench_csrcsc.py --bits 8 > $OUT/csrcsc_bench.txt 2>&1; cat $OUT/csrcsc_bench.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"radix_|scan_|segment_" -c 20 --csv --log-file $OUT/csrcsc_launches.csv python tools/bench_csrcsc.py --bits 8 --iters 1 > $OUT/ncu_stdout.txt 2>&1
python tools/launch_list.py $OUT/csrcsc_launches.csv --per-launch kernel > $OUT/csrcsc_per_launch.txt 2>&1; head -16 $OUT/csrcsc_per_launch.txt
( time timeout 1500 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_t15/bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "x_of_bound", d["e2e"].get("x_of_bound"))
        print("roofline", {k: d["roofline"].get(k) for k in ("frac", "frac_hybrid", "frac_3xtf32", "alt_frac_3xtf32", "own_peaks_tflops", "own_peaks_error")})
        print("launches", d["gpu_launches"], d.get("gpu_launches_detail"), "cpu", d["cpu_baseline"])
        for k, v in d["extra"].items():
            print(k, json.dumps({kk: v.get(kk) for kk in ("value", "unit", "ms", "ms_per_iter", "bench_seconds", "error", "trace")})[:300])
            if k == "e2e_file": print("   ", json.dumps(v)[:900])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
