#!/bin/bash
# round 2, call A (1 GPU): parity tests, the three bench configs (cfg4 / cfg5 reduced N as a baseline of the general path),
# then one ncu --set full capture of the fused feature kernel (Phi-only, BASELINE configs[1]).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err; cut -c1-1500 gpurun_out/bench_cfg3.json
python bench.py --config cfg4 --n 10e6 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4_10M.json 2> gpurun_out/bench_cfg4.err; tail -3 gpurun_out/bench_cfg4.err
python bench.py --config cfg5 --n 5e6 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg5_5M.json 2> gpurun_out/bench_cfg5.err; tail -3 gpurun_out/bench_cfg5.err
python tools/feat_bench.py 2e6 > gpurun_out/feat_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"features2" -s 1 -c 1 -o /tmp/prof_feat python tools/feat_bench.py 2e6 > gpurun_out/ncu_feat.log 2>&1
tail -2 gpurun_out/ncu_feat.log; cat gpurun_out/feat_plain.log
ncu -i /tmp/prof_feat.ncu-rep --page raw --csv > gpurun_out/prof_feat_raw.csv 2>/dev/null
ncu -i /tmp/prof_feat.ncu-rep --page source --csv > gpurun_out/prof_feat_src.csv 2>/dev/null
ls -la /tmp/prof_feat.ncu-rep; SZ=$(stat -c %s /tmp/prof_feat.ncu-rep); if [ $SZ -lt 30000000 ]; then cp /tmp/prof_feat.ncu-rep gpurun_out/; fi
