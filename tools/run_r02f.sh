#!/bin/bash
# round 2, call F (1 GPU): parity tests incl. features VJP and device full-covariance KL; f64 headline still at its number
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python tools/kernel_breakdown.py 8 11 30 1 gaussian 4e6 2>&1 | tail -1
