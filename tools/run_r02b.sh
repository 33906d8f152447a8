#!/bin/bash
# round 2, call B (1 GPU): parity tests with the new general fused backward kernel, then per-kernel breakdowns
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for args in "8 11 30 1 gaussian 4e6" "8 11 30 1 gaussian 4e6 no_fused_bwd=1" "8 13 30 1 gaussian 4e6" "8 13 30 1 gaussian 4e6 no_bwd4=1" \
            "32 5 64 1 bernoulli 4e6" "32 5 64 1 bernoulli 4e6 no_bwd4=1" "8 11 64 1 gaussian 2e6" "8 11 64 1 gaussian 2e6 no_bwd4=1" "8 11 30 1 bernoulli 4e6"; do
  python tools/kernel_breakdown.py $args 2>&1 | tail -1 | tee -a gpurun_out/breakdown.log
done
python bench.py --steps 5 --no-extras > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err; cut -c1-900 gpurun_out/bench_cfg3.json
