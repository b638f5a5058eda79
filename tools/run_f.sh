#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/fog_share.py < /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fog_share_launches.csv python tools/fog_share.py > /dev/null 2>&1 < /dev/null
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/fog_share_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]:
    print(r[4][:40].ljust(42), r[8], r[-1])
PY
