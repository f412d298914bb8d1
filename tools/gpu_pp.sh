#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/pingpong > gpurun_out/pingpong.log 2>&1
