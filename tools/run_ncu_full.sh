#!/bin/bash
# ncu capture of per-chunk kernels of ONE timed step; exports the raw page as CSV (small) and keeps
# the .ncu-rep only if it is below 40 MB (gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out
N=${1:-227328}; KREGEX=${2:-"zonal_bwd|zonal_fwd"}; COUNT=${3:-2}; TAG=${4:-r01c}; SKIP=${5:-0}
python tools/quick_bench.py $N 30 1 > gpurun_out/plain.log 2>&1 && \
GPFY_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"$KREGEX" -s $SKIP -c $COUNT -o /tmp/prof_$TAG python tools/quick_bench.py $N 30 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
for i in $(seq 0 $((COUNT-1))); do ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof_${TAG}_src$i.csv 2>/dev/null; done
ls -la /tmp/prof_$TAG.ncu-rep gpurun_out/
SZ=$(stat -c %s /tmp/prof_$TAG.ncu-rep); if [ $SZ -lt 40000000 ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
