#!/bin/bash
# features3_kernel: A/B against features2_kernel, then the feature-related parity tests
mkdir -p gpurun_out
timeout 300 python tools/feat3_check.py 1e7 2>&1 | tail -7 | tee gpurun_out/feat3.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_feat3.log
