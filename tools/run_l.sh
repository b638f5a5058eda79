#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q -k "sched or hygiene or long_rays or config4" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err < /dev/null
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_final2_n1.json').read().strip().splitlines()[-1])
for k,o in (('c4',j),('c2',j['c2']),('c3',j['c3_fog'])):
    print(k, o['value'], o['ms_per_step'], o['e2e']['value'], o['roofline']['frac'], o.get('without_tile_cost_history'), o['cpu_baseline']['parity']['mismatched_pixels'])
PY
