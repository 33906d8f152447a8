#!/bin/bash
# features3_kernel: A/B against features2_kernel
mkdir -p gpurun_out
timeout 300 python tools/feat3_check.py 1e7 2>&1 | tail -7 | tee gpurun_out/feat3.log
