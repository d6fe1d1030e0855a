#!/bin/bash
# 2 GPUs: f-AnoGAN with the fused peer optimiser (equivalence + c5 lines), e2e with pinned map fetches
TAG=${1:-r2y}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|DP_EQUIV|Error|rror" gpurun_out/${TAG}_pytest.log | tail -12
run2() { # name, extra env, port, bench args
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --no-cpu-baseline $4 2> gpurun_out/${TAG}_$1.err | grep '^{' > gpurun_out/${TAG}_$1.json
  cut -c1-260 gpurun_out/${TAG}_$1.json; grep -iE "error|trap|fail" gpurun_out/${TAG}_$1.err | head -5
}
run2 c5_peer "UAD_PEER_ADAM=1" 29546 "--config c5 --steps 15 --warmup 3"
run2 c5_nccl "UAD_PEER_ADAM=0" 29547 "--config c5 --steps 15 --warmup 3"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench1.json').read()); print(d['ms_per_step'], d['value'], d['e2e'])"
