#!/bin/bash
# Runs the FP64 peak microbenchmark with clock sampling; writes gpurun_out/fp64_peak.json + clocks csv.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/fp64_peak_clocks.csv &
SMI=$!
./tools/fp64_peak > gpurun_out/fp64_peak.json
RC=$?
kill $SMI
nproc > gpurun_out/host_cores.txt; lscpu | head -20 >> gpurun_out/host_cores.txt
cat gpurun_out/fp64_peak.json
exit $RC
