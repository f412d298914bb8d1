#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/final_check.log 2>&1
timeout 100 python -m pytest tests/test_gpu_chol.py tests/test_gpu_gbmv.py -m gpu -x -q 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
