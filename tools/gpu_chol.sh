#!/bin/bash
# Cholesky pass: parity tests + timings
mkdir -p gpurun_out
exec > gpurun_out/chol.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python tools/time_chol.py 131072 1024 U 4
timeout 200 python tools/time_chol.py 131072 1024 L 1
timeout 200 python tools/time_chol.py 131072 256 U 1
cat > /tmp/np.py <<'PY'
import sys
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
bm.handle(0).tune("pb_nopdl", 1)
sys.argv = ["x", "131072", "1024", "U", "1"]
exec(open("tools/time_chol.py").read())
PY
python /tmp/np.py
