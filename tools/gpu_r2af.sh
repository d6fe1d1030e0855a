#!/bin/bash
# last call of the round: GPU suite on the committed tree + the single-pass mode at 64 slices per GPU
TAG=${1:-r2af}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -4
timeout 120 python bench.py --config c4 --batch 64 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_b64.json 2> gpurun_out/${TAG}_bench_c4_b64.err
cut -c1-260 gpurun_out/${TAG}_bench_c4_b64.json
