#!/bin/bash
# ncu --set full capture of the level-set kernel on c4 (the dense instantiation)
cd "$(dirname "$0")/.."
rm -f gpurun_out/r02f_*.ncu-rep
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_levelset -s 3 -c 1 -o gpurun_out/r02f_c4_ls python tools/prof_c4.py 5 > gpurun_out/prof_c4.log 2>&1 < /dev/null
ls -la gpurun_out/r02f*
