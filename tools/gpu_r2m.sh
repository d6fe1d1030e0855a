#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|AssertionError|Error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-420 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
