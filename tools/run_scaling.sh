#!/bin/bash
mkdir -p gpurun_out
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -c 300 gpurun_out/bench_n$n.err | tail -2
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print($n, round(d['value']/1e6,2), 'M pts/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']/1e6,2))
PY
done
