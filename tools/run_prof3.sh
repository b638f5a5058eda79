#!/bin/bash
# final ncu captures: level-set kernel on c4 and c2 (history-ordered frames), fog shadow + primary on c3, launch list of the bench command
cd "$(dirname "$0")/.."
rm -f gpurun_out/r02e_*.ncu-rep
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_levelset -s 3 -c 1 -o gpurun_out/r02e_c4_ls python tools/prof_c4.py 5 > gpurun_out/prof_c4.log 2>&1 < /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_levelset -s 3 -c 1 -o gpurun_out/r02e_c2_ls python tools/prof_c2.py 5 device > gpurun_out/prof_c2.log 2>&1 < /dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fog_shadow -s 2 -c 1 -o gpurun_out/r02e_c3_shadow python tools/fog_ab.py c3 > /dev/null 2>&1 < /dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fog_primary -s 2 -c 1 -o gpurun_out/r02e_c3_primary python tools/fog_ab.py c3 > /dev/null 2>&1 < /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1 < /dev/null
ls -la gpurun_out/r02e*
