#!/bin/bash
# zonal_fwd3_kernel: A/B against zonal_fwd2_kernel
mkdir -p gpurun_out
timeout 300 python tools/fwd3_check.py 2>&1 | tail -7 | tee gpurun_out/fwd3.log
