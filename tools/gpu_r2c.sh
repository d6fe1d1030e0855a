#!/bin/bash
# round-2 GPU call c: where conv_halo_ss spends its time (debug-switch matrix + one ncu --set full capture per layer shape)
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 300 python tools/time_hs.py > gpurun_out/${TAG}_time_hs.txt 2>&1
cat gpurun_out/${TAG}_time_hs.txt
MODES=0 REPS=1 WARM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_ss -o gpurun_out/${TAG}_hs python tools/time_hs.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/
