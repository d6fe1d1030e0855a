#!/bin/bash
# end-of-round-2 evidence: full GPU suite, bench lines (c2 with cpu baseline + layer table, c4, c5, reference arm), ncu launch list of one
# eager step, ncu --set full of its tcgen05 launches (3xTF32 at B = 64 and 1xTF32 at B = 32), sanitizer passes, other configurations
TAG=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -6
timeout 400 python bench.py --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config c4 --layer-table gpurun_out/${TAG}_layers_c4.json > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-300 gpurun_out/${TAG}_bench_c4.json; tail -2 gpurun_out/${TAG}_bench_c4.err
timeout 300 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
cut -c1-300 gpurun_out/${TAG}_bench_c5.json; tail -2 gpurun_out/${TAG}_bench_c5.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cut -c1-300 gpurun_out/${TAG}_bench_ref.json
# ncu: side stream off so that the launch order is the layer order
export UAD_SIDE_WGRAD=0 UAD_DENSE_FORK=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_ss|wgrad_ss" --launch-skip 27 -c 27 -f -o gpurun_out/${TAG}_tc3 python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_tc3.log 2>&1
tail -2 gpurun_out/${TAG}_tc3.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_ss|wgrad_ss" --launch-skip 27 -c 27 -f -o gpurun_out/${TAG}_tc1 python tools/profile_step.py 2 tc1 32 > gpurun_out/${TAG}_tc1.log 2>&1
tail -2 gpurun_out/${TAG}_tc1.log
unset UAD_SIDE_WGRAD UAD_DENSE_FORK
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv2d_fwd_dgrad_wgrad or convT2d_fwd_dgrad_wgrad" ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -4
( time B=2 MODES=0 CASES=2,3,7 REPS=1 WARM=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/time_hs.py ) > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_racecheck.log | tail -4
timeout 300 python tools/config_times.py > gpurun_out/${TAG}_config_times.txt 2>&1; tail -8 gpurun_out/${TAG}_config_times.txt
