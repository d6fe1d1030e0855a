#!/bin/bash
# first hardware run of UAD_MATH_TC_1XTF32: op parity (modes 0 / 1 / 2), the 1xTF32 step test, kernel times, c4 / c2 bench lines
TAG=${1:-r2p}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider -s -k "conv2d_fwd_dgrad_wgrad or convT2d_fwd_dgrad_wgrad or 1xtf32" ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|1xTF32|Error|error" gpurun_out/${TAG}_pytest.log | tail -12
MATH=2 MODES=0,2,27 timeout 300 python tools/time_hs.py > gpurun_out/${TAG}_time_hs_tc1.txt 2>&1; cat gpurun_out/${TAG}_time_hs_tc1.txt
timeout 300 python bench.py --config c4 --steps 200 --warmup 5 --layer-table gpurun_out/${TAG}_layers_c4.json > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 300 python bench.py --config c4 --batch 64 --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_c4_b64.json > gpurun_out/${TAG}_bench_c4_b64.json 2> gpurun_out/${TAG}_bench_c4_b64.err
cut -c1-400 gpurun_out/${TAG}_bench_c4_b64.json; tail -3 gpurun_out/${TAG}_bench_c4_b64.err
