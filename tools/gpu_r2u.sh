#!/bin/bash
# 2 GPUs: NCCL equivalence incl. the bucketed in-graph all-reduce; c2 / c4 bench lines with the decoder bucket overlapped vs the
# collective behind the replay; MC-dropout evaluation test
TAG=${1:-r2u}
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_golden.py -m gpu -q -x -p no:cacheprovider -s -k "two_gpu or monte" ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|DP_EQUIV|Error|error" gpurun_out/${TAG}_pytest.log | tail -10
run2() { # name, extra env, bench args
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline $4 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  cut -c1-230 gpurun_out/${TAG}_$1.json; tail -2 gpurun_out/${TAG}_$1.err | cut -c1-300
}
run2 c2_buckets "UAD_DP_BUCKETS=1" 29544 ""
run2 c2_outside "UAD_DP_BUCKETS=0" 29545 ""
run2 c4_buckets "UAD_DP_BUCKETS=1" 29546 "--config c4"
run2 c4_outside "UAD_DP_BUCKETS=0" 29547 "--config c4"
