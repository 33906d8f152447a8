#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_i8.log
for prec in f64 f64i8; do
GPFY_BENCH_PRECISION=$prec timeout 300 python tools/kernel_breakdown.py 8 11 30 1 gaussian 4e6 2>&1 | tail -1 | tee -a gpurun_out/breakdown_i8.log
done
