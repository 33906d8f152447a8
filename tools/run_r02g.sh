#!/bin/bash
# round 2, call G (8 GPUs): BASELINE configs at their stated scale - cfg3 (N = 10 M), cfg4 (N = 100 M, 13 degrees), cfg5 (N = 20 M,
# d = 32, Bernoulli), each as one torchrun job over 8 ranks with the sharded-vs-single / two-device self-checks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for cfg in cfg3 cfg4 cfg5; do
  timeout 600 $TR --master-port 2951${#cfg} bench.py --gpus 8 --config $cfg --steps 5 --warmup 3 > gpurun_out/bench_${cfg}_n8.json 2> gpurun_out/bench_${cfg}_n8.err
  echo "== $cfg rc=$?"; tail -2 gpurun_out/bench_${cfg}_n8.err | cut -c1-300
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${cfg}_n8.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","elbo_data_term","checks")}, "e2e", d["e2e"]["value"], "step frac", d["roofline"]["step"]["frac_of_fp64_peak"])
except Exception as e: print("no line:", e)
P
done
timeout 600 $TR --master-port 29520 bench.py --gpus 8 --config cfg3 --dtype f32 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_f32_n8.json 2> gpurun_out/bench_cfg3_f32_n8.err; echo "== f32 rc=$?"; cut -c1-200 gpurun_out/bench_cfg3_f32_n8.json
