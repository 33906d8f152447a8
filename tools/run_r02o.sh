#!/bin/bash
# Re-entry baseline at the restored state: GPU tests, headline bench, per-kernel breakdown of the BASELINE shapes.
mkdir -p gpurun_out; rm -f gpurun_out/breakdown.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for args in "8 11 30 1 gaussian 4e6" "8 13 30 1 gaussian 4e6" "32 5 64 1 bernoulli 4e6"; do
  python tools/kernel_breakdown.py $args 2>&1 | tail -1 | tee -a gpurun_out/breakdown.log
done
python tools/feat_bench.py 1e7 2>&1 | tail -2 | tee -a gpurun_out/breakdown.log
