#!/bin/bash
# FIRST GPU call of round 2 (~15-20 GPU-minutes; every stage has its own timeout, worst case ~55 min): answers the open hardware questions left at the end of round 1, which
# had no GPU minutes left when the candidates were written.  usage: gpurun --timeout 3400 -- 'bash tools/gpu_round2_entry.sh'
#   1. the shipped defaults (even stage ring at N = 64, N = 128 on the first-generation kernel) were derived on CPU from the
#      barrier-protocol model - confirm the GPU suite and re-measure the headline;
#   2. tools/ubench/operand_probe.cu: tf32 operand forms the Form-W redesign needs (MN-major SWIZZLE_128B_BASE32B operands,
#      row-shifted descriptors, truncation of raw fp32 words);
#   3. the N = 32 candidate kernel gather_gemm_tc3 (UAD_TC_V3=1): correctness, then time against the shipped kernel;
#   4. the compositions that so far only ran through the CPU emulation of the ABI (AnoVAEGAN, AAE / constrained AAE, CE, GMVAE incl. its latent kernel pair):
#      real-kernel parity, CUDA-graph replay, trainers (UAD_UNVERIFIED=1) - drop the skip markers of the files that pass;
#   5. swizzled epilogue staging (UAD_TC_V2=21): N = 128 column-split dual issue on an even four-stage ring;
#   6. plane-resident Form-W kernel wgrad_tc2 (UAD_WGRAD_V2=1) - read experiment E7 of step 2 first;
#   7. SS-form gather kernel gather_gemm_ss (UAD_TC_SS bit mask) - pre-split activations, no converter warps;
#   8. halo-resident N = 32 kernel gather_gemm_tc_np_halo (UAD_TC_HALO=1) - a k-block loads only its weight image.
TAG=${1:-r2a}
mkdir -p gpurun_out build
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/operand_probe tools/ubench/operand_probe.cu \
  && timeout 120 build/operand_probe > gpurun_out/${TAG}_operand_probe.txt 2>&1
cat gpurun_out/${TAG}_operand_probe.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/l2_to_sm tools/ubench/l2_to_sm.cu \
  && timeout 120 build/l2_to_sm > gpurun_out/${TAG}_l2_to_sm.txt 2>&1      # the L2 -> shared-memory ceiling the gather kernels run into next
cat gpurun_out/${TAG}_l2_to_sm.txt
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
UAD_TC_V3=1 timeout 300 python -m pytest tests/test_gpu_v3_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_v3_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_v3_pytest.log
timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_shipped.txt 2>&1
UAD_TC_V3=1 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_v3.txt 2>&1
tail -12 gpurun_out/${TAG}_time_tc_shipped.txt gpurun_out/${TAG}_time_tc_v3.txt
UAD_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_anovaegan.py tests/test_gpu_aae.py tests/test_gpu_ce.py tests/test_gpu_gmvae.py -m gpu -q --maxfail=20 \
  -p no:cacheprovider > gpurun_out/${TAG}_unverified_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_unverified_pytest.log
UAD_TC_V2=5 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_v2_5.txt 2>&1      # N = 128 column-split (odd ring: timing only)
UAD_TC_V2=0 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_v1.txt 2>&1        # first-generation kernel everywhere
tail -6 gpurun_out/${TAG}_time_tc_v2_5.txt gpurun_out/${TAG}_time_tc_v1.txt
# 5. swizzled epilogue staging (gather_gemm_tc2_swz): four stages at N = 128 -> column-split dual issue on an EVEN ring, six at N = 64
UAD_TC_V2=21 timeout 300 python -m pytest tests/test_gpu_swz_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_swz_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_swz_pytest.log
UAD_TC_V2=21 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_v2_21.txt 2>&1
UAD_TC_V2=17 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_v2_17.txt 2>&1    # swizzled staging, N = 64 only (6 stages)
tail -6 gpurun_out/${TAG}_time_tc_v2_21.txt gpurun_out/${TAG}_time_tc_v2_17.txt
# 6. plane-resident Form-W candidate (wgrad_tc2): only meaningful if experiment E7 of the operand probe passed (see step 2's output)
UAD_WGRAD_V2=1 timeout 300 python -m pytest tests/test_gpu_wgrad_v2_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_wgrad2_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_wgrad2_pytest.log
UAD_WGRAD_V2=1 timeout 300 python bench.py --steps 30 --warmup 5 --layer-table gpurun_out/${TAG}_layers_wgrad2.json > gpurun_out/${TAG}_bench_wgrad2.json 2> gpurun_out/${TAG}_bench_wgrad2.err
cat gpurun_out/${TAG}_bench_wgrad2.json
# 7. SS-form candidate (gather_gemm_ss): pre-split activations, TMA -> MMA ring without converters
UAD_TC_SS=7 timeout 300 python -m pytest tests/test_gpu_ss_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_ss_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_ss_pytest.log
UAD_TC_SS=15 timeout 300 python -m pytest tests/test_gpu_ss_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_ss_rawhi_pytest.log 2>&1   # raw tensor as hi operand: passes iff the tensor core truncates (E1)
tail -3 gpurun_out/${TAG}_ss_rawhi_pytest.log
for m in 1 3 7 9; do
  UAD_TC_SS=$m timeout 300 python bench.py --steps 30 --warmup 5 --layer-table gpurun_out/${TAG}_layers_ss$m.json > gpurun_out/${TAG}_bench_ss$m.json 2> gpurun_out/${TAG}_bench_ss$m.err
  cat gpurun_out/${TAG}_bench_ss$m.json
done
# per-kernel event timings of the two newest candidates (column 0 of each line = the kernel as it would ship)
UAD_WGRAD_V2=1 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_wgrad2.txt 2>&1
UAD_TC_SS=7 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_ss7.txt 2>&1
tail -11 gpurun_out/${TAG}_time_tc_wgrad2.txt gpurun_out/${TAG}_time_tc_ss7.txt
# one-screen summary LAST, so that it is what gpurun's tail shows
python tools/summarize_round2.py ${TAG} > gpurun_out/${TAG}_summary.txt 2>&1
cat gpurun_out/${TAG}_summary.txt
# 8. halo-resident N = 32 candidate (gather_gemm_tc_np_halo, stride-1 form): a k-block moves 8 KB instead of 24 KB
UAD_TC_HALO=1 timeout 300 python -m pytest tests/test_gpu_halo_candidate.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_halo_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_halo_pytest.log
UAD_TC_HALO=1 timeout 200 python tools/time_tc.py > gpurun_out/${TAG}_time_tc_halo.txt 2>&1
UAD_TC_HALO=1 timeout 300 python bench.py --steps 30 --warmup 5 --layer-table gpurun_out/${TAG}_layers_halo.json > gpurun_out/${TAG}_bench_halo.json 2> gpurun_out/${TAG}_bench_halo.err
python tools/summarize_round2.py ${TAG} > gpurun_out/${TAG}_summary.txt 2>&1
cat gpurun_out/${TAG}_summary.txt
