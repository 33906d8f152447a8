#!/bin/bash
# round 2, call C (1 GPU): tcgen05 descriptor probe, parity tests, per-kernel breakdowns after the SYRK split / GEMM padding skip
mkdir -p gpurun_out
timeout 60 ./tools/umma_probe 2>&1 | tee gpurun_out/umma_probe.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/breakdown.log
for args in "8 11 30 1 gaussian 4e6" "8 13 30 1 gaussian 4e6" "32 5 64 1 bernoulli 4e6" "8 11 64 1 gaussian 2e6" "8 11 100 1 gaussian 1e6" "2 11 1000 1 gaussian 4e6"; do
  python tools/kernel_breakdown.py $args 2>&1 | tail -1 | tee -a gpurun_out/breakdown.log
done
