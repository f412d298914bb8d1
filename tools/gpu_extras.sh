#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/extras.log 2>&1
timeout 600 python -m pytest tests/test_gpu_typed.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
import bench_extras
r = bench_extras.run_secondary(bm)
print(json.dumps({k: r[k] for k in ("TYPED", "GBMV_wide")}))
PY
