#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/${TAG}_ops_pytest.log 2>&1
tail -30 gpurun_out/${TAG}_ops_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
a=json.load(open('gpurun_out/r2h_layers.json'))
for r in a['table']:
    if 'wgrad' in r['op']: print(r['op'], round(r['ms'],4))
PY
