#!/bin/bash
# 2 GPUs, final code: data-parallel equivalence test (incl. pooled ROC) + default bench line at 2 ranks
TAG=${1:-r2ad}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|DP_EQUIV|Error|rror" gpurun_out/${TAG}_pytest.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 200 --warmup 5 2> gpurun_out/${TAG}_bench2.err | grep '^{' > gpurun_out/${TAG}_bench2.json
cut -c1-260 gpurun_out/${TAG}_bench2.json; grep -iE "error|trap|fail" gpurun_out/${TAG}_bench2.err | head -5
