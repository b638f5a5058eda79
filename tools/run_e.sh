#!/bin/bash
cd "$(dirname "$0")/.."
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t4.log 2>&1 < /dev/null
tail -5 gpurun_out/t4.log
(time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err < /dev/null
tail -3 gpurun_out/bench_r2b_n1.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_r2b_n1.json").read().strip().splitlines()[-1])
print("c4", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l["cpu_baseline"]["value"], l["cpu_baseline"]["parity"])
for k in ("c2","c3_fog"):
    print(k, l[k]["value"], l[k]["ms_per_step"], l[k]["e2e"]["value"], l[k]["roofline"]["frac"], l[k]["cpu_baseline"]["value"], l[k]["cpu_baseline"]["parity"], l[k]["gpu_launches_per_step"])
print(json.dumps(l.get("extras"))[:1500])
PY
