#!/bin/bash
# end-of-round-2 evidence: full GPU suite, bench lines (c2 with cpu baseline + layer table, c4), ncu launch list of one eager step, ncu
# metrics of its tcgen05 launches (3xTF32 at B = 64, 1xTF32 at B = 32), one --set full report of the dominant kernel (kept as .ncu-rep),
# sanitizer passes, other configurations.  gpurun copies back at most 64 MiB: the big reports are digested on the box and deleted.
TAG=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -6
timeout 400 python bench.py --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read()); print('clocks', d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['with_maps']['value'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'loss_check', d['loss_check'])"
timeout 300 python bench.py --config c4 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_c4.json > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-300 gpurun_out/${TAG}_bench_c4.json; tail -2 gpurun_out/${TAG}_bench_c4.err
# ncu: side stream off so that the launch order is the layer order
export UAD_SIDE_WGRAD=0 UAD_DENSE_FORK=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for MODE in "tc3 64" "tc1 32"; do
  set -- $MODE
  timeout 600 ncu --metrics $M --clock-control none -k "regex:conv_halo_ss|wgrad_ss" --launch-skip 27 -c 27 -f -o gpurun_out/${TAG}_$1 python tools/profile_step.py 2 $1 $2 > gpurun_out/${TAG}_$1.log 2>&1
  tail -1 gpurun_out/${TAG}_$1.log
  python tools/ncu_summary.py gpurun_out/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_summary.json
  rm -f gpurun_out/${TAG}_$1.ncu-rep
done
# (the --set full report of the two largest launches is taken by tools/gpu_r2aa.sh: -k regex:conv_halo_ss|wgrad_ss --launch-skip 36 -c 2)
unset UAD_SIDE_WGRAD UAD_DENSE_FORK
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv2d_fwd_dgrad_wgrad or convT2d_fwd_dgrad_wgrad" ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -4
( time B=2 MODES=0 CASES=2,3,7 REPS=1 WARM=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/time_hs.py ) > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_racecheck.log | tail -4
timeout 300 python tools/config_times.py > gpurun_out/${TAG}_config_times.txt 2>&1; tail -30 gpurun_out/${TAG}_config_times.txt | grep -E "config|ms_per_step|slices"
du -sh gpurun_out
