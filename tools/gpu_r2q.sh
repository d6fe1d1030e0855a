#!/bin/bash
# ncu --set full with source-level sampling of the two N = 32 conv_halo_ss launches (T4.fwd / T4.dgrad), 3xTF32 and 1xTF32
TAG=${1:-r2q}
mkdir -p gpurun_out
for M in 1 2; do
MATH=$M CASES=0,1 REPS=1 WARM=1 MODES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_ss -c 4 -f -o gpurun_out/${TAG}_hs_math$M python tools/time_hs.py > gpurun_out/${TAG}_ncu_math$M.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_math$M.log
done
ls -la gpurun_out/${TAG}*
