#!/bin/bash
# ncu --set full of the int8 contraction (one 0.9 M-point chunk) with source-level stall samples
mkdir -p gpurun_out
export GPFY_BENCH_PRECISION=f64i8
python tools/kernel_breakdown.py 8 11 30 1 gaussian 909312 > gpurun_out/plain_i8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_i8" -s 2 -c 1 -o /tmp/prof_i8 python tools/kernel_breakdown.py 8 11 30 1 gaussian 909312 > gpurun_out/ncu_i8.log 2>&1
tail -1 gpurun_out/ncu_i8.log
ncu -i /tmp/prof_i8.ncu-rep --page raw --csv > gpurun_out/prof_i8_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8.ncu-rep --page source --csv > gpurun_out/prof_i8_src.csv 2>/dev/null
ls -la /tmp/prof_i8.ncu-rep
