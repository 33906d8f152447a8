#!/bin/bash
# ncu --set full of the f64 GEMM (with the psi epilogue) and SYRK at the 13-degree shape (general path)
mkdir -p gpurun_out
python tools/kernel_breakdown.py 8 13 30 1 gaussian 1818624 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_dmma|syrk_dmma" -s 4 -c 2 -o /tmp/prof_g13 python tools/kernel_breakdown.py 8 13 30 1 gaussian 1818624 > gpurun_out/ncu_g13.log 2>&1
tail -1 gpurun_out/ncu_g13.log
ncu -i /tmp/prof_g13.ncu-rep --page raw --csv > gpurun_out/prof_g13_raw.csv 2>/dev/null
for i in 0 1; do ncu -i /tmp/prof_g13.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof_g13_src$i.csv 2>/dev/null; done
