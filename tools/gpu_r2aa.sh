#!/bin/bash
# compile-time slicing flag: parity + bench; --set full reports (with source) of the dominant kernel and of dec_Conv2DT_4's input gradient
TAG=${1:-r2aa}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -4
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config c4 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-300 gpurun_out/${TAG}_bench_c4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aa_layers.json'))
for r in d['table'][:12]: print('  ', r['op'], round(r['ms'],4))
PY
export UAD_SIDE_WGRAD=0 UAD_DENSE_FORK=0
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_ss|wgrad_ss" --launch-skip 36 -c 2 -f -o gpurun_out/${TAG}_dec4_wgrad_dgrad_full python tools/profile_step.py 2 tc3 > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log; ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
