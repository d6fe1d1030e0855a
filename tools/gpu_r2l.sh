#!/bin/bash
TAG=${1:-r2l}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|AssertionError|Error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 5 --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "test_conv2d_fwd_dgrad_wgrad or test_convT" ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|deselected" gpurun_out/${TAG}_memcheck.log | tail -5
