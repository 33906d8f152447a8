#!/bin/bash
# full GPU suite + headline bench (int8 contraction default) + DMMA-only bench for comparison
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --dtype f64dmma --no-e2e --no-cpu-baseline --no-extras > gpurun_out/bench_dmma.json 2> gpurun_out/bench_dmma.err; tail -3 gpurun_out/bench_dmma.err; cut -c1-400 gpurun_out/bench_dmma.json
