#!/bin/bash
# epilogue rewrite (constants after the transpose) + 1xTF32 without a lo pass / with doubled stages: parity, kernel times, bench lines
TAG=${1:-r2r}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|1xTF32|Error|error" gpurun_out/${TAG}_pytest.log | tail -12
MATH=1 MODES=0,2,27 timeout 300 python tools/time_hs.py > gpurun_out/${TAG}_time_hs_tc3.txt 2>&1; cat gpurun_out/${TAG}_time_hs_tc3.txt
MATH=2 MODES=0,2,27 timeout 300 python tools/time_hs.py > gpurun_out/${TAG}_time_hs_tc1.txt 2>&1; cat gpurun_out/${TAG}_time_hs_tc1.txt
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-420 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config c4 --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_c4.json > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-420 gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 300 python bench.py --config c4 --batch 64 --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_c4_b64.json > gpurun_out/${TAG}_bench_c4_b64.json 2> gpurun_out/${TAG}_bench_c4_b64.err
cut -c1-420 gpurun_out/${TAG}_bench_c4_b64.json; tail -3 gpurun_out/${TAG}_bench_c4_b64.err
