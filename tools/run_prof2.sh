#!/bin/bash
# ncu --set full captures of the level-set kernel (history-ordered frames): c4 and c2
cd "$(dirname "$0")/.."
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_levelset -s 3 -c 1 -o gpurun_out/r02c_c4_ls python tools/prof_c4.py 5 > gpurun_out/prof_c4.log 2>&1 < /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_levelset -s 3 -c 1 -o gpurun_out/r02c_c2_ls python tools/prof_c2.py 5 device > gpurun_out/prof_c2.log 2>&1 < /dev/null
ls -la gpurun_out/r02c*.ncu-rep
