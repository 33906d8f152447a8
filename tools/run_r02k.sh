#!/bin/bash
# ncu --set full of the tcgen05 weighted-Gram kernel only
mkdir -p gpurun_out
python tools/f32_step.py 1818624 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"syrk_tf32x3" -s 2 -c 1 -o /tmp/prof_s32 python tools/f32_step.py 1818624 > gpurun_out/ncu_s32.log 2>&1
tail -1 gpurun_out/ncu_s32.log
ncu -i /tmp/prof_s32.ncu-rep --page raw --csv > gpurun_out/prof_s32_raw.csv 2>/dev/null
ncu -i /tmp/prof_s32.ncu-rep --page source --csv > gpurun_out/prof_s32_src.csv 2>/dev/null
