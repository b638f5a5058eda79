#!/bin/bash
cd "$(dirname "$0")/.."
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/t5.log 2>&1 < /dev/null
tail -4 gpurun_out/t5.log
timeout 300 python __graft_entry__.py smoke < /dev/null 2>&1 | tail -2
(time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/bench_final_ref_n1.json 2> gpurun_out/bench_final_ref_n1.err < /dev/null
(time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err < /dev/null
tail -3 gpurun_out/bench_final_n1.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_final_ref_n1.json").read().strip().splitlines()[-1])
l=json.loads(open("gpurun_out/bench_final_n1.json").read().strip().splitlines()[-1])
print("same config:", r["config"] == l["config"], "ref", r["value"])
print("c4", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l["roofline"]["issue_slot_util"], l["without_tile_cost_history"], l["cpu_baseline"]["value"], l["cpu_baseline"]["parity"])
for k in ("c2","c3_fog"):
    print(k, l[k]["value"], l[k]["ms_per_step"], l[k]["e2e"]["value"], l[k]["roofline"]["frac"], l[k]["roofline"]["traffic"], l[k].get("without_tile_cost_history"), l[k]["cpu_baseline"]["value"], l[k]["cpu_baseline"]["parity"], l[k]["gpu_launches_per_step"])
print(json.dumps(l.get("extras"))[:700])
PY
