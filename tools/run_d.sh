#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fog_shadow -s 2 -c 1 -o gpurun_out/r02_c3_shadow python tools/fog_ab.py c3 > gpurun_out/prof_shadow.log 2>&1 < /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fog_primary -s 2 -c 1 -o gpurun_out/r02_c3_primary python tools/fog_ab.py c3 > gpurun_out/prof_primary.log 2>&1 < /dev/null
ls -la gpurun_out/*.ncu-rep
