#!/bin/bash
# side-stream filter gradients + forked dense backward: full GPU suite, bench lines (c2, c2 without the side stream, c4, c5)
TAG=${1:-r2s}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -8
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
UAD_SIDE_WGRAD=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_noside.json 2> gpurun_out/${TAG}_bench_noside.err
cut -c1-300 gpurun_out/${TAG}_bench_noside.json; tail -3 gpurun_out/${TAG}_bench_noside.err
UAD_DENSE_FORK=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_nofork.json 2> gpurun_out/${TAG}_bench_nofork.err
cut -c1-300 gpurun_out/${TAG}_bench_nofork.json; tail -3 gpurun_out/${TAG}_bench_nofork.err
timeout 300 python bench.py --config c4 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-300 gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 600 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
cat gpurun_out/${TAG}_bench_c5.json; tail -5 gpurun_out/${TAG}_bench_c5.err
