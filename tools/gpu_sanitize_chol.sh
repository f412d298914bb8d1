#!/bin/bash
# compute-sanitizer racecheck over one small blocked Cholesky + solve case; prints which kernels it flags
mkdir -p gpurun_out
exec > gpurun_out/sanitize_chol.log 2>&1
timeout 100 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 3000 python -m pytest tests/test_gpu_chol.py -m gpu -x -q -k "wide_blocked and shape0" > /tmp/racecheck_raw.log 2>&1
echo "racecheck rc=$?"
grep -c "hazard detected" /tmp/racecheck_raw.log
grep -A2 "hazard detected" /tmp/racecheck_raw.log | grep -o " in [A-Za-z_0-9<>,: ]*(" | sort | uniq -c | sort -rn | head -12
grep "hazard detected" /tmp/racecheck_raw.log | grep -o "hazard detected (.*" | sort | uniq -c | sort -rn | head -5
grep -m1 -A12 "hazard detected" /tmp/racecheck_raw.log
tail -3 /tmp/racecheck_raw.log
