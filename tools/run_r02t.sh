#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/breakdown_t.log
for prec in f64dmma f64; do
for args in "8 13 30 1 gaussian 4e6" "32 5 64 1 bernoulli 4e6" "8 11 64 1 gaussian 2e6" "2 11 12 1 gaussian 4e6"; do
  GPFY_BENCH_PRECISION=$prec python tools/kernel_breakdown.py $args 2>&1 | tail -1 | tee -a gpurun_out/breakdown_t.log
done; done
