#!/bin/bash
# 2 GPUs: fused peer-memory optimiser step (reduce-scatter + Adam + all-gather) vs NCCL all-reduce + Adam
TAG=${1:-r2v}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|DP_EQUIV|Error|error|rror" gpurun_out/${TAG}_pytest.log | tail -12
run2() { # name, extra env, port, bench args
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline $4 2> gpurun_out/${TAG}_$1.err | grep '^{' > gpurun_out/${TAG}_$1.json
  cut -c1-230 gpurun_out/${TAG}_$1.json; grep -iE "error|trap|fail" gpurun_out/${TAG}_$1.err | head -5
}
run2 c2_peer "UAD_PEER_ADAM=1" 29544 ""
run2 c2_nccl "UAD_PEER_ADAM=0" 29545 ""
run2 c4_peer "UAD_PEER_ADAM=1" 29546 "--config c4"
run2 c4_nccl "UAD_PEER_ADAM=0" 29547 "--config c4"
