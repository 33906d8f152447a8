#!/bin/bash
# ncu --set full of the per-chunk kernels on the cfg4 shape (13 degrees, M = 340: zonal_bwd4 path, dynamic weighted-Gram units, int8 contraction with 5 out-blocks)
mkdir -p gpurun_out
python tools/kernel_breakdown.py 8 13 30 1 gaussian 909312 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"syrk_dmma|gemm_i8|zonal_fwd2|zonal_bwd4|lik_point" -s 10 -c 5 -o /tmp/prof_c4 python tools/kernel_breakdown.py 8 13 30 1 gaussian 909312 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log
ncu -i /tmp/prof_c4.ncu-rep --page raw --csv > gpurun_out/prof_c4_raw.csv 2>/dev/null
