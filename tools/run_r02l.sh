#!/bin/bash
# f32 path: drain period sweep of the tcgen05 weighted-Gram kernel (speed and accuracy)
mkdir -p gpurun_out; rm -f gpurun_out/breakdown_f32.log
timeout 300 python -m pytest tests/test_gpu_f32.py -x -q 2>&1 | tail -3
for kf in 8 32 128 512; do
  GPFY_BENCH_PRECISION=f32 python tools/kernel_breakdown.py 8 11 30 1 gaussian 4e6 syrk32_kf=$kf 2>&1 | tail -1 | tee -a gpurun_out/breakdown_f32.log
done
GPFY_BENCH_PRECISION=f32 python tools/kernel_breakdown.py 32 5 64 1 bernoulli 4e6 2>&1 | tail -1 | tee -a gpurun_out/breakdown_f32.log
GPFY_BENCH_PRECISION=f32 python tools/kernel_breakdown.py 8 13 30 1 gaussian 4e6 2>&1 | tail -1 | tee -a gpurun_out/breakdown_f32.log
