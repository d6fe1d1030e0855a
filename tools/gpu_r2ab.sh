#!/bin/bash
# Cin = 1 kernels rewritten (multi-row blocks, 16-byte shared-memory reads): parity, smoke(), bench
TAG=${1:-r2ab}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ab_layers.json'))
for r in d['table']:
    if 'enc_conv2D_0' in r['op']: print('  ', r['op'], round(r['ms'],4), round(r['frac'] or 0,3))
PY
