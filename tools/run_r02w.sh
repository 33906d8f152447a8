#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_cfg3_n$N.err | tail -n 1 > gpurun_out/bench_cfg3_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n$N.json')); print($N, round(d['value']/1e6,1), 'M pts/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), d['elbo_data_term'])"
