#!/bin/bash
# round 2, call I (1 GPU): ncu --set full of the two tcgen05 kernels of the f32 path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_f32.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_f32.log
python tools/f32_step.py 1818624 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tf32x3" -s 4 -c 2 -o /tmp/prof_f32 python tools/f32_step.py 1818624 > gpurun_out/ncu_f32.log 2>&1
tail -2 gpurun_out/ncu_f32.log
ncu -i /tmp/prof_f32.ncu-rep --page raw --csv > gpurun_out/prof_f32_raw.csv 2>/dev/null
for i in 0 1; do ncu -i /tmp/prof_f32.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof_f32_src$i.csv 2>/dev/null; done
ls -la /tmp/prof_f32.ncu-rep
