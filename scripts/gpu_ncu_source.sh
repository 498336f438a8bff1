#!/bin/bash
# ON THE GPU BOX: ncu --set full with source correlation on ONE kernel family, export the per-line
# stall samples as CSV (top lines only). usage: gpu_ncu_source.sh <tag> <kernel regex> [skip] [count]
set -u
TAG=$1; RE=$2; SKIP=${3:-0}; CNT=${4:-1}
CMD="python scripts/layer_profile.py"
export CHUNK=512
$CMD > gpurun_out/src_plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/src_plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${RE}" \
    -s $SKIP -c $CNT -o /tmp/src_${TAG} $CMD > gpurun_out/src_ncu_${TAG}.log 2>&1
ncu -i /tmp/src_${TAG}.ncu-rep --page source --csv --print-source sass > /tmp/src_${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/src_${TAG}.ncu-rep --page source --csv --print-source cuda > /tmp/src_${TAG}_cuda.csv 2>/dev/null
ncu -i /tmp/src_${TAG}.ncu-rep --page raw --csv > gpurun_out/src_${TAG}_raw.csv 2>/dev/null
head -c 3000000 /tmp/src_${TAG}_cuda.csv > gpurun_out/src_${TAG}_cuda.csv
head -c 6000000 /tmp/src_${TAG}_sass.csv > gpurun_out/src_${TAG}_sass.csv
tail -3 gpurun_out/src_ncu_${TAG}.log; ls -la gpurun_out/src_${TAG}*
