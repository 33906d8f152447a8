#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/quick_bench.py 1e7 30 3 2>&1 | tee gpurun_out/quick.log
python tools/quick_bench.py 227328 30 2 > gpurun_out/quick_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_quick.csv python tools/quick_bench.py 227328 30 1 > gpurun_out/ncu_quick.log 2>&1
tail -3 gpurun_out/ncu_quick.log
