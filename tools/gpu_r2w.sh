#!/bin/bash
# 8 GPUs: the configs BASELINE.json names for 8 x B200 - c4 (32 slices per GPU) with the fused peer optimiser and with NCCL, c5 (f-AnoGAN, 16 per GPU)
TAG=${1:-r2w}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run8() { # name, extra env, port, bench args
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --no-cpu-baseline $4 2> gpurun_out/${TAG}_$1.err | grep '^{' > gpurun_out/${TAG}_$1.json
  cut -c1-230 gpurun_out/${TAG}_$1.json; grep -iE "error|trap|fail" gpurun_out/${TAG}_$1.err | head -5
}
run8 c4_peer "UAD_PEER_ADAM=1" 29546 "--config c4 --steps 300 --warmup 5"
run8 c4_nccl "UAD_PEER_ADAM=0" 29547 "--config c4 --steps 300 --warmup 5"
run8 c5 "UAD_PEER_ADAM=1" 29548 "--config c5 --steps 15 --warmup 3"
