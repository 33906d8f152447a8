#!/bin/bash
# last check of the library as rebuilt at the committed source: smoke and the parity files around the forward / feature kernels
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 300 python -m pytest tests/test_gpu_barrier_free.py tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_last.log
