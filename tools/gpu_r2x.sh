#!/bin/bash
# image-pair layout for the 8 x 8 M-grids: op parity (all modes), step parity, bench c2 / c4
TAG=${1:-r2x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py tests/test_gpu_golden.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -8
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config c4 --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers_c4.json > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cut -c1-300 gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
python - <<'PY'
import json
for f in ['gpurun_out/TAG_layers.json','gpurun_out/TAG_layers_c4.json']:
    d=json.load(open(f.replace('TAG','r2x')))
    print(f, round(d['ms_per_step_graph'],4))
    for r in d['table']:
        if 'enc_conv2D_4' in r['op'] or 'dec_Conv2DT_0' in r['op']: print('  ',r['op'], round(r['ms'],4))
PY
