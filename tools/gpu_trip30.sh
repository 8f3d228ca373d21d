#!/bin/bash
# Trip 30: final-state check on one GPU: full GPU suite, smoke, bench (both arms), ncu launch list of the bench command.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary30.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; t0=$SECONDS; timeout ${TMO:-1200} "$@" > $OUT/$name.log 2>&1; echo "exit $? wall $((SECONDS - t0)) s" | tee -a $S; tail -n ${TAILN:-6} $OUT/$name.log | cut -c1-3500 | tee -a $S; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')"
run bench_ref python bench.py --impl reference
run bench python bench.py
grep "^{" $OUT/bench.log > $OUT/bench.json
run launches_bench ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra
python tools/launch_list.py $OUT/launches_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu --no-extra (32768^3, trip 30)" > $OUT/launches_bench_trip30.txt
head -12 $OUT/launches_bench_trip30.txt | tee -a $S
rm -f $OUT/launches_bench.csv
