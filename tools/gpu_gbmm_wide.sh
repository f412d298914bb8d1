#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/gbmm_wide.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_gbmm.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/time_gbmm.py 65536 1024
timeout 200 python tools/time_gbmm.py 262144 256
timeout 200 python tools/time_gbmm.py 1048576 128
timeout 200 python tools/time_gbmm.py 1048576 64
cat > /tmp/w.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
hd = bm.handle(0); hd.tune("gbmm_wide", 1)
sys.argv = ["x", "1048576", "64"]
exec(open("tools/time_gbmm.py").read())
sys.argv = ["x", "4194304", "32"]
exec(open("tools/time_gbmm.py").read())
PY
timeout 200 python /tmp/w.py
