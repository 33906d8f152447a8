#!/bin/bash
mkdir -p gpurun_out
B=${1:-tools/i8_gemm_bench}
$B 288 909312 1 > gpurun_out/plain_i8b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_i8" -s 1 -c 1 -o /tmp/prof_i8b $B 288 909312 1 > gpurun_out/ncu_i8b.log 2>&1
tail -1 gpurun_out/ncu_i8b.log
ncu -i /tmp/prof_i8b.ncu-rep --page raw --csv > gpurun_out/prof_i8b_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8b.ncu-rep --page source --csv > gpurun_out/prof_i8b_src.csv 2>/dev/null
