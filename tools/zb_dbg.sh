#!/bin/bash
for d in 0 1; do
echo "zf dbg=$d"; GPFY_ZF_DBG=$d python bench.py --no-cpu-baseline --no-e2e --steps 2 --n 4546560 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['kernels_ms_per_step'].items()})"
done
