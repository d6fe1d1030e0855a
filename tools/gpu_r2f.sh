#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=10 -p no:cacheprovider ) > gpurun_out/${TAG}_ops_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_ops_pytest.log
MODES=0,2,27 timeout 300 python tools/time_hs.py > gpurun_out/${TAG}_time_hs.txt 2>&1
cat gpurun_out/${TAG}_time_hs.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
