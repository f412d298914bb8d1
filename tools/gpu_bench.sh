#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/bench_run.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["unit"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k, v in d["configs"].items():
    print(k, json.dumps({kk: vv for kk, vv in v.items() if kk not in ("roofline", "cpu_baseline", "parity", "workload", "cholesky")})[:400])
    print("   parity", json.dumps(v.get("parity"))[:500])
    if "cholesky" in v:
        print("   cholesky", json.dumps(v["cholesky"])[:1500])
PY
