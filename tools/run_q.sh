#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/quant_ab.py < /dev/null 2>&1 | tail -8
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_render_levelset --csv --log-file gpurun_out/quant_dram.csv python tools/quant_ab.py > /dev/null 2>&1 < /dev/null
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/quant_dram.csv')) if len(r)>10 and r[0].isdigit()]
# 3 metrics per launch; 6 launches per configuration, 5 configurations
per={}
for r in rows:
    per.setdefault(r[0],{})[r[-3]]=(r[-1], r[-2], r[4][:70])
ids=sorted(per, key=int)
for k in range(0, len(ids), 6):
    last=per[ids[min(k+5, len(ids)-1)]]
    print(last.get('dram__bytes_read.sum'), last.get('dram__bytes_write.sum'), last.get('gpu__time_duration.sum'))
PY
