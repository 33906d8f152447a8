#!/bin/bash
# parity + short bench (development loop)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 3 ${1:-} 2>gpurun_out/quick.err | tail -1 > gpurun_out/quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick.json'))
print(f"{d['value']/1e6:.2f} M pts/s  {d['ms_per_step']:.2f} ms  frac {d['roofline']['step']['frac_of_fp64_peak']:.4f}  e2e {d['e2e']['value']/1e6 if d['e2e'] else 0:.2f}")
print({k: round(v,2) for k,v in d['roofline']['kernels_ms_per_step'].items()}, d['clocks'])
PY
