#!/bin/bash
# timeline of the blocked Cholesky's kernels from the stamped debug build (make EXTRA=-DPB_DEBUG_TIMING -> tools/_dbg_libbmb200.so)
mkdir -p gpurun_out
exec > gpurun_out/chol_dbg.log 2>&1
cp bandedmatrices.jl_b200/libbmb200.so /tmp/rel.so
cp tools/_dbg_libbmb200.so bandedmatrices.jl_b200/libbmb200.so
timeout 200 python tools/time_chol.py 4096 1024 U 1 | grep STAMP | tail -28
cp /tmp/rel.so bandedmatrices.jl_b200/libbmb200.so
