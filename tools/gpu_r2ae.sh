#!/bin/bash
# 8 GPUs, final kernels: c4 (32 slices per GPU, fused peer optimiser) - refreshes the scaling table of DESIGN 6
TAG=${1:-r2ae}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --config c4 --steps 300 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_c4_peer.err | grep '^{' > gpurun_out/${TAG}_c4_peer.json
cut -c1-260 gpurun_out/${TAG}_c4_peer.json; grep -iE "error|trap|fail" gpurun_out/${TAG}_c4_peer.err | head -5
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --config c4 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_c4_1gpu.json 2> gpurun_out/${TAG}_c4_1gpu.err
cut -c1-260 gpurun_out/${TAG}_c4_1gpu.json
