#!/bin/bash
mkdir -p gpurun_out
python tools/kernel_breakdown.py 8 11 30 1 gaussian 909312 no_fused_bwd=1 > gpurun_out/plain_b4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"zonal_bwd4" -s 2 -c 1 -o /tmp/prof_b4 python tools/kernel_breakdown.py 8 11 30 1 gaussian 909312 no_fused_bwd=1 > gpurun_out/ncu_b4.log 2>&1
tail -1 gpurun_out/ncu_b4.log
ncu -i /tmp/prof_b4.ncu-rep --page raw --csv > gpurun_out/prof_b4_raw.csv 2>/dev/null
ncu -i /tmp/prof_b4.ncu-rep --page source --csv > gpurun_out/prof_b4_src.csv 2>/dev/null
