#!/bin/bash
# zonal_fwd3_kernel with degree-by-degree row tiles: A/B against the aligned walk and zonal_fwd2_kernel, then the parity files that reach it
mkdir -p gpurun_out
timeout 200 python tools/fwd3_check.py 2>&1 | tail -6 | tee gpurun_out/fwd3b.log
timeout 300 python -m pytest tests/test_gpu_barrier_free.py tests/test_gpu_f32.py tests/test_gpu_i8.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_fwd3b.log
