#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/quick_bench.py 340992 30 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_dmma|syrk_dmma|lik_bwd|zonal_fwd|dvhat" -s 39 -c 5 -o gpurun_out/prof_r01a python tools/quick_bench.py 340992 30 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
