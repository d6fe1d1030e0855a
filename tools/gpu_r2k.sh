#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests/test_gpu_step.py -m gpu -q --maxfail=25 -p no:cacheprovider -k benched -s ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|AssertionError|worst gradient|rel-err" gpurun_out/${TAG}_pytest.log | tail -12
