#!/bin/bash
# r02 trip 18b: the full GPU suite died with a fatal signal in trip 18a and only the tail of the log was kept.
# Same command with -v, full log; then the test files after the one that crashed, in a second process.
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_t18; mkdir -p $OUT
timeout 150 python -X faulthandler -m pytest tests -m gpu -v -p no:cacheprovider > $OUT/tests_v.txt 2>&1; echo "pytest rc=$?" >> $OUT/tests_v.txt
grep -v "PASSED\|SKIPPED" $OUT/tests_v.txt | head -60
python - <<'PY'
import re, subprocess, sys, glob
log = open("gpurun_out/r02_t18/tests_v.txt").read()
if "Fatal Python error" not in log and "rc=0" in log.splitlines()[-1]:
    sys.exit(0)
names = re.findall(r"^(tests/test_gpu_\w+\.py)::", log, re.M)
if not names: sys.exit(0)
crash_file = names[-1]
rest = [f for f in sorted(glob.glob("tests/test_gpu_*.py")) if f > crash_file]
print("crashed in", crash_file, "-> running", rest, flush=True)
if rest:
    r = subprocess.run([sys.executable, "-m", "pytest", *rest, "-m", "gpu", "-q", "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=100)
    open("gpurun_out/r02_t18/tests_rest.txt", "w").write(r.stdout + r.stderr)
    print((r.stdout + r.stderr)[-1500:])
PY
