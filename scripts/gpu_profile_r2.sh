#!/bin/bash
# ON THE GPU BOX (gpurun): round-2 ncu launch lists (time, DRAM bytes, tensor-pipe activity) of every
# librotmv kernel of ONE warm inference forward (configs[1]) and ONE warm training step (configs[3]).
# The scripts bracket their third pass with cudaProfilerStart/Stop, so the capture is exactly one
# pass whatever the number of kernels per C-ABI call. Plain runs first (a number printed under ncu
# is never a bench value).
set -u
TAG=${1:-r2}
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
python scripts/one_forward.py > gpurun_out/${TAG}_fwd_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_fwd_plain.log; exit 1; }
python scripts/one_train_step.py > gpurun_out/${TAG}_train_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_train_plain.log; exit 1; }
timeout 600 ncu --metrics $MET --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_fwd_launches.csv python scripts/one_forward.py > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 900 ncu --metrics $MET --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_train_launches.csv python scripts/one_train_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
cat gpurun_out/${TAG}_fwd_plain.log gpurun_out/${TAG}_train_plain.log
wc -l gpurun_out/${TAG}_fwd_launches.csv gpurun_out/${TAG}_train_launches.csv
