#!/bin/bash
TAG=${1:-r2ac}
mkdir -p gpurun_out
for W in 4 3 6 8 12; do UAD_WS_WAVES=$W timeout 120 python tools/time_wgrad.py; done 2>&1 | tee gpurun_out/${TAG}_wgrad_waves.txt
for D in 2048 8192 16384; do UAD_WS_DEPTH=$D timeout 120 python tools/time_wgrad.py; done 2>&1 | tee -a gpurun_out/${TAG}_wgrad_waves.txt
