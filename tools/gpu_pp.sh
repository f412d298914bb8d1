#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/dmma_lat > gpurun_out/dmma_lat.log 2>&1
