#!/bin/bash
# 2-GPU call: NCCL data-parallel equivalence (incl. the in-graph all-reduce) and a 2-rank bench line beside the 1-rank one
TAG=${1:-r2n}
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -p no:cacheprovider -s ) > gpurun_out/${TAG}_dp_pytest.log 2>&1
grep -E "passed|failed|DP_EQUIV|Error" gpurun_out/${TAG}_dp_pytest.log | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
cut -c1-260 gpurun_out/${TAG}_bench2.json; tail -2 gpurun_out/${TAG}_bench2.err
UAD_GRAPH_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench2_outside.json 2> gpurun_out/${TAG}_bench2_outside.err
cut -c1-260 gpurun_out/${TAG}_bench2_outside.json
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
cut -c1-260 gpurun_out/${TAG}_bench1.json
