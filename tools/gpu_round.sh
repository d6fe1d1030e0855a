#!/bin/bash
# One GPU-box session: full GPU test suite, headline bench (+ per-kernel table), f-AnoGAN and restoration timings.
# Everything lands in gpurun_out/.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
for args in "256 16 1 10 0" "256 16 1 10 1" "256 128 1 4 1"; do
  timeout 200 python tools/fanogan_time.py $args >> gpurun_out/${TAG}_fanogan.jsonl 2>> gpurun_out/${TAG}_fanogan.err
done
cat gpurun_out/${TAG}_fanogan.jsonl
timeout 200 python tools/restore_time.py 256 110 150 1 > gpurun_out/${TAG}_restore.json 2> gpurun_out/${TAG}_restore.err
cat gpurun_out/${TAG}_restore.json
timeout 300 python tools/config_times.py 20 > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
cat gpurun_out/${TAG}_configs.json
if [ "${2:-}" = "ncu" ]; then
  # launch list of one eager step (the second of two) and a --set full capture of the launches selected by $3 (regex)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 128 -c 128 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_ncu1.log 2>&1
  KRE=${3:-gather_gemm_tc|wgrad_tc}
  NK=${4:-27}
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" -s ${NK} -c ${NK} \
      -o gpurun_out/${TAG}_tc python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_ncu2.log 2>&1
  ncu -i gpurun_out/${TAG}_tc.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_raw.csv 2>/dev/null
  ls -la gpurun_out/ | tail -20
fi
