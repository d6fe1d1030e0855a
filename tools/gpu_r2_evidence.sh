#!/bin/bash
# round-2 evidence run: ncu launch list of one eager train step, ncu --set full of the tcgen05 launches of a step, sanitizer passes,
# the other configurations' step times.  Files land in gpurun_out/r2e2_*; the digests are copied to profiles/ by hand.
TAG=${1:-r2e2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_ss|wgrad_ss" --launch-skip 23 -c 23 -o gpurun_out/${TAG}_tc python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_tc.log 2>&1
tail -2 gpurun_out/${TAG}_tc.log
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv and tc" ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/${TAG}_memcheck.log | tail -5
( time B=1 MODES=0 CASES=2,3,7 REPS=1 WARM=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/time_hs.py ) > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_racecheck.log | tail -5
timeout 300 python tools/config_times.py > gpurun_out/${TAG}_config_times.txt 2>&1; tail -8 gpurun_out/${TAG}_config_times.txt
timeout 600 python tools/fanogan_time.py > gpurun_out/${TAG}_fanogan_times.txt 2>&1; tail -8 gpurun_out/${TAG}_fanogan_times.txt
