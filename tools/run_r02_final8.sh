#!/bin/bash
# Round-2 final 8-GPU lines: cfg3 (N = 10 M), cfg4 (N = 100 M, 13 degrees), cfg5 (N = 20 M, d = 32, Bernoulli), int8 contraction default
mkdir -p gpurun_out
for cfg in cfg3 cfg4 cfg5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config $cfg --steps 5 --warmup 3 > gpurun_out/bench_${cfg}_n8.json 2> gpurun_out/bench_${cfg}_n8.err
tail -c 300 gpurun_out/bench_${cfg}_n8.err | tail -2
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${cfg}_n8.json').read().strip().splitlines()[-1])
print('$cfg', round(d['value']/1e6,2), 'M pts/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']/1e6,2), d['checks'], d['elbo_data_term'])
PY
done
