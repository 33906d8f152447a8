#!/bin/bash
# round 2, call H (1 GPU): first run of the tcgen05 weighted-Gram kernel (f32 path v2) + the re-shaped host slab ramp
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_f32.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_f32.log
timeout 300 python bench.py --dtype f32 --steps 5 --no-extras > gpurun_out/bench_cfg3_f32.json 2> gpurun_out/bench_cfg3_f32.err; tail -3 gpurun_out/bench_cfg3_f32.err
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg3_f32.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"], d["roofline"]["frac"], {k:(v["max_rel"] if isinstance(v,dict) else v) for k,v in d.get("f32_vs_f64").items()})
except Exception as e: print("no line", e)
P
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_f32.py 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --no-extras --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_cfg3.json").read().strip().splitlines()[-1])
print("f64", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"])
P
