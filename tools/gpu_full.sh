#!/bin/bash
# full GPU pass: parity tests, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k, v in d["configs"].items():
    print(k, {kk: vv for kk, vv in v.items() if kk.endswith("_ms") or kk.endswith("_us") or kk in ("ms", "lu_ms_incl_widen", "solve_ms", "GFLOPs", "lu_TFLOPs")}, v.get("parity"))
PY
# ncu summary of the slot-scheduled solve (forward = 2nd-last, backward = last gbtrs_slot launch of the case)
timeout 300 ncu --clock-control none --set full -k regex:gbtrs_slot -s 2 -c 2 -o gpurun_out/p_slot -f python tools/prof_case.py lu 131072 > /dev/null 2>&1
ncu -i gpurun_out/p_slot.ncu-rep --page raw --csv > gpurun_out/slot_r2_raw.csv 2>/dev/null; rm -f gpurun_out/p_slot.ncu-rep
