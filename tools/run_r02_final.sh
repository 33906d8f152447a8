#!/bin/bash
# Round-2 final evidence on one B200: GPU tests, smoke, both bench arms, the DMMA-only and f32 lines, the ncu launch list of the
# bench command and one `ncu --set full` capture of the four per-chunk kernels on one 0.9 M-point chunk.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
python bench.py --dtype f64dmma --no-cpu-baseline --no-extras > gpurun_out/bench_dmma.json 2> gpurun_out/bench_dmma.err; cut -c1-200 gpurun_out/bench_dmma.json
python bench.py --dtype f32 --no-cpu-baseline --no-extras > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; cut -c1-200 gpurun_out/bench_f32.json
for cfg in cfg4 cfg5; do
python bench.py --config $cfg --n 4e6 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/bench_${cfg}_4M.json 2> gpurun_out/bench_${cfg}.err; cut -c1-200 gpurun_out/bench_${cfg}_4M.json
done
python bench.py --n 1818624 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --n 1818624 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tools/quick_bench.py 909312 30 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"syrk_dmma|gemm_i8|zonal_fwd2|zonal_bwd3" -s 8 -c 4 -o /tmp/prof_final python tools/quick_bench.py 909312 30 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
for i in 0 1 2 3; do ncu -i /tmp/prof_final.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof_final_src$i.csv 2>/dev/null; done
ls -la gpurun_out | head -40
