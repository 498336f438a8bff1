#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain bench, ncu launch list, one ncu --set full capture of the
# dominant kernel. Only text summaries (+ a small .ncu-rep) are left under gpurun_out/.
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --chunk 512"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
$CMD > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 150 -c 10 \
    -o /tmp/igemm_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/igemm_${TAG}.ncu-rep --page raw --csv > gpurun_out/igemm_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/igemm_${TAG}.ncu-rep --page details --csv > gpurun_out/igemm_${TAG}_details.csv 2>/dev/null
SZ=$(stat -c %s /tmp/igemm_${TAG}.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 0 ] && [ "$SZ" -lt 30000000 ]; then cp /tmp/igemm_${TAG}.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -12
