#!/bin/bash
# N-GPU check: the scaling bench line the driver will run (default flags), N = $1 (default 8)
N=${1:-8}
mkdir -p gpurun_out
exec > gpurun_out/multi$N.log 2>&1
set -x
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e", d.get("e2e", {}).get("value"))
print("sharded_check", d.get("sharded_check"))
for k, v in d.get("configs", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("ms", "gbtrf_ms", "gbtrs_ms", "sharding", "sharded_bit_identical", "error", "wall_s")}, v.get("parity"))
PY
tail -3 gpurun_out/bench_n$N.err
