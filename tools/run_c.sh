#!/bin/bash
# per-kernel times of the fog paths on C3 (ncu launch list), then the NanoVDB prior-art bar on c2 / c4
cd "$(dirname "$0")/.."
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fog_launches.csv python tools/fog_ab.py c3 > /dev/null 2>&1 < /dev/null
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/fog_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:40].ljust(42), r[8], r[-1])
PY
timeout 300 python tools/nanovdb_bar.py c2 c4 < /dev/null
