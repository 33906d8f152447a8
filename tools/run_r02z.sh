#!/bin/bash
# 2 B200: the two-device tests and the headline bench under torchrun at the closing state (barrier-free forward in the sharded path)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_cfg3_n2.err | tail -n 1 > gpurun_out/bench_cfg3_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n2.json')); print(2, round(d['value']/1e6,1), 'M pts/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), d['elbo_data_term'], d.get('checks'))"
