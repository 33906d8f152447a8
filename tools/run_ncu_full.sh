#!/bin/bash
# one full-set ncu capture of the per-chunk kernels of ONE timed step (N = 2 exact chunks)
mkdir -p gpurun_out
N=${1:-227328}
python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -2
python tools/quick_bench.py $N 30 1 > gpurun_out/plain.log 2>&1 && \
GPFY_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"syrk_dmma|zonal_bwd|zonal_fwd|dvhat" -c 4 -o gpurun_out/prof_r01c python tools/quick_bench.py $N 30 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
