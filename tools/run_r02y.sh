#!/bin/bash
# Round-2 closing evidence after the barrier-free forward / feature kernels: GPU tests, smoke, the headline bench line, one
# `ncu --set full` capture of features3_kernel and zonal_fwd3_kernel, the ncu launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
timeout 200 python tools/feat_bench.py 2e6 > gpurun_out/feat_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"features3|zonal_fwd3" -s 2 -c 2 -o /tmp/prof_bf python tools/feat_bench.py 2e6 > gpurun_out/ncu_bf.log 2>&1
tail -1 gpurun_out/ncu_bf.log; cat gpurun_out/feat_plain.log
ncu -i /tmp/prof_bf.ncu-rep --page raw --csv > gpurun_out/prof_bf_raw.csv 2>/dev/null
for i in 0 1; do ncu -i /tmp/prof_bf.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof_bf_src$i.csv 2>/dev/null; done
python bench.py --n 1818624 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --n 1818624 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
