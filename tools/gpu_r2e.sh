#!/bin/bash
# round-2 GPU call e: source-level sampling of conv_halo_ss (two N = 32 shapes; full mode and protocol-only mode)
TAG=${1:-r2e}
mkdir -p gpurun_out
for m in 0 27; do
CASES=0,1 MODES=$m REPS=1 WARM=0 timeout 600 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --clock-control none --import-source on -k regex:conv_halo_ss -o gpurun_out/${TAG}_hs_m$m python tools/time_hs.py > gpurun_out/${TAG}_ncu_m$m.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_m$m.log
done
