#!/bin/bash
# round 2, call E (1 GPU): first run of the tcgen05 f32 path - its tests under a short timeout, then timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_f32.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_f32.log
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_f32.py 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --dtype f32 --steps 5 --no-extras --no-e2e > gpurun_out/bench_cfg3_f32.json 2> gpurun_out/bench_cfg3_f32.err; tail -3 gpurun_out/bench_cfg3_f32.err; cut -c1-300 gpurun_out/bench_cfg3_f32.json
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg3_f32.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["roofline"]["frac"], d.get("f32_vs_f64"))
except Exception as e: print("no line", e)
P
