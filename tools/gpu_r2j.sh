#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python tools/wgrad_err.py 2>&1 | tail -4
( time timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_ops.py -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|AssertionError|worst gradient" gpurun_out/${TAG}_pytest.log | tail -12
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-330 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
