#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/c4chol.log 2>&1
python - <<'PY'
import sys, json, ctypes
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
import bench_configs
print(json.dumps(bench_configs.run_c4_cholesky(bm, object()), indent=1))
PY
