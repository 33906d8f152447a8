#!/bin/bash
# zonal_fwd3_kernel: one ncu --set full capture on one 0.9 M-point chunk (after the plain run of the same command exited 0)
mkdir -p gpurun_out
timeout 200 python tools/quick_bench.py 909312 30 1 > gpurun_out/plain2.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"zonal_fwd3" -s 2 -c 1 -o /tmp/prof_fwd3 python tools/quick_bench.py 909312 30 1 > gpurun_out/ncu_fwd3.log 2>&1
tail -1 gpurun_out/ncu_fwd3.log; tail -1 gpurun_out/plain2.log
ncu -i /tmp/prof_fwd3.ncu-rep --page raw --csv > gpurun_out/prof_fwd3_raw.csv 2>/dev/null
ncu -i /tmp/prof_fwd3.ncu-rep --page source --csv > gpurun_out/prof_fwd3_src.csv 2>/dev/null
