#!/bin/bash
# round 2, call D (2 GPUs): tcgen05 MN-major descriptor candidates, the 2-GPU tests pytest skips on one GPU, and the bench under
# torchrun with its sharded-vs-single and one-process-two-devices self-checks
mkdir -p gpurun_out
timeout 60 ./tools/umma_probe 2>&1 | tee gpurun_out/umma_probe.log
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err
tail -3 gpurun_out/bench_cfg3_n2.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_cfg3_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","elbo_data_term","checks")}, d["e2e"]["value"], d["checksum"])
P
