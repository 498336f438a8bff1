#!/bin/bash
# ON THE GPU BOX: ncu --set full with source correlation on one kernel of scripts/bnconv_probe.py.
# usage: gpu_ncu_probe.sh <tag> <kernel regex> [skip] [count]
set -u
TAG=$1; RE=$2; SKIP=${3:-2}; CNT=${4:-1}
CMD="python scripts/bnconv_probe.py"
$CMD > gpurun_out/probe_plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/probe_plain_${TAG}.log; exit 1; }
cat gpurun_out/probe_plain_${TAG}.log
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${RE}" \
    -s $SKIP -c $CNT -o /tmp/probe_${TAG} $CMD > gpurun_out/probe_ncu_${TAG}.log 2>&1
ncu -i /tmp/probe_${TAG}.ncu-rep --page source --csv --print-source sass > /tmp/probe_${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/probe_${TAG}.ncu-rep --page raw --csv > gpurun_out/probe_${TAG}_raw.csv 2>/dev/null
head -c 6000000 /tmp/probe_${TAG}_sass.csv > gpurun_out/probe_${TAG}_sass.csv
tail -2 gpurun_out/probe_ncu_${TAG}.log
