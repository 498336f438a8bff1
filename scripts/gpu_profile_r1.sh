#!/bin/bash
# ON THE GPU BOX (gpurun): launch list + DRAM/tensor counters of every kernel of one warm forward
# (configs[1]: B=256, V=2, bf16), and one --set full capture of representative igemm launches.
set -u
TAG=${1:-r1b}
CMD="python scripts/one_forward.py"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
NL=$(grep -m1 -o '[0-9]* launches' gpurun_out/${TAG}_plain.log | cut -d' ' -f1)
echo "launches per forward: $NL"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s $((2 * NL)) -c $NL --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
# full sections for the first 24 igemm launches of the warm forward (layer1 + layer2 shapes)
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 126 -c 63 \
    -o /tmp/${TAG}_igemm $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_igemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_igemm_raw.csv 2>/dev/null
SZ=$(stat -c %s /tmp/${TAG}_igemm.ncu-rep 2>/dev/null || echo 0)
echo "ncu-rep size $SZ"
if [ "$SZ" -gt 0 ] && [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_igemm.ncu-rep gpurun_out/; fi
tail -3 gpurun_out/${TAG}_plain.log
