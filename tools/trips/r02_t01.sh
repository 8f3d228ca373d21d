#!/bin/bash
# r02 trip 1: full GPU suite (incl. parity against the reference's own binaries, oracle/_ref) + box facts.
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t01; mkdir -p $OUT
{ nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi --query-gpu=name,memory.total --format=csv; df -h /dev/shm /tmp | cat; uname -r; } > $OUT/box.txt 2>&1
ls oracle/_ref | head -50 >> $OUT/box.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=15 2>&1 | tail -40 > $OUT/tests.txt
tail -5 $OUT/tests.txt
