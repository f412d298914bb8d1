#!/bin/bash
# full GPU pass: every parity test, smoke
mkdir -p gpurun_out
exec > gpurun_out/full.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3
