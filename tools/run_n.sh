#!/bin/bash
# usage: tools/run_n.sh N  -- both bench arms under torchrun on N GPUs of this box
cd "$(dirname "$0")/.."
N=${1:-2}
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err < /dev/null
tail -4 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_r2_n$N.json").read().strip().splitlines()[-1])
print("c4 N=$N", l["value"], l["ms_per_step"], "kernel", l["kernel_ms"], "sync", l["sync_ms"], "e2e", l["e2e"]["value"], l["e2e"]["frame_matches_device_path"], l["gpu_launches"])
for k in ("c2","c3_fog"):
    print(k, l[k]["value"], l[k]["ms_per_step"], "kernel", l[k]["kernel_ms"], "e2e", l[k]["e2e"]["value"], l[k]["e2e"]["frame_matches_device_path"], l[k]["gpu_launches_per_step"])
PY
