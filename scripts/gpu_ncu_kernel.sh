#!/bin/bash
# ON THE GPU BOX: ncu --set full of a kernel regex inside a command; exports raw + sass-source CSV.
# usage: gpu_ncu_kernel.sh <tag> <regex> <skip> <count> <command...>
set -u
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
"$@" > gpurun_out/k_plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/k_plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${RE}" \
    -s $SKIP -c $CNT -o /tmp/k_${TAG} "$@" > gpurun_out/k_ncu_${TAG}.log 2>&1
ncu -i /tmp/k_${TAG}.ncu-rep --page raw --csv > gpurun_out/k_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/k_${TAG}.ncu-rep --page source --csv --print-source sass > /tmp/k_${TAG}_sass.csv 2>/dev/null
head -c 4000000 /tmp/k_${TAG}_sass.csv > gpurun_out/k_${TAG}_sass.csv
tail -2 gpurun_out/k_ncu_${TAG}.log
