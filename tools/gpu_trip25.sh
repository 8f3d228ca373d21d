#!/bin/bash
# Trip 25: full GPU suite after the CallGuard / quiesce change, resident bench at the final panel width.
set -u
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary25.txt
: > $S
run() { name=$1; shift; echo "=== $name" | tee -a $S; timeout ${TMO:-900} "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $S; tail -n ${TAILN:-8} $OUT/$name.log | cut -c1-3000 | tee -a $S; }
run tests python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run resident_bench python tools/resident_bench.py
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')"
