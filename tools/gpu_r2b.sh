#!/bin/bash
# round-2 GPU call b: first hardware run of conv_halo_ss (default) beside the converter-warp kernels (UAD_HS=0)
TAG=${1:-r2b}
mkdir -p gpurun_out build
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/operand_probe tools/ubench/operand_probe.cu \
  && timeout 180 build/operand_probe > gpurun_out/${TAG}_operand_probe.txt 2>&1
tail -40 gpurun_out/${TAG}_operand_probe.txt
( time timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=10 -p no:cacheprovider ) > gpurun_out/${TAG}_ops_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_ops_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_hs1.json > gpurun_out/${TAG}_bench_hs1.json 2> gpurun_out/${TAG}_bench_hs1.err
cat gpurun_out/${TAG}_bench_hs1.json; tail -3 gpurun_out/${TAG}_bench_hs1.err
UAD_HS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_hs0.json > gpurun_out/${TAG}_bench_hs0.json 2> gpurun_out/${TAG}_bench_hs0.err
cat gpurun_out/${TAG}_bench_hs0.json; tail -3 gpurun_out/${TAG}_bench_hs0.err
