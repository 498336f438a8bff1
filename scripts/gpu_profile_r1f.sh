#!/bin/bash
# ON THE GPU BOX (gpurun): round-1 final (r1f) ncu launch lists: every librotmv kernel of one warm
# inference forward, and the kernels that changed since r1e (bn_apply, head_loss, rotate_gather) of
# one warm training step. Time + DRAM bytes + tensor-pipe activity per launch.
set -u
TAG=r1f
MINE='regex:igemm_|wgrad_|stem_|maxpool|avgpool|rotate_gather|head_loss|bn_|relu_bwd|colsum|permute_cast|adam_|dilate|simt_'
CHANGED='regex:bn_apply|head_loss|rotate_gather'
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
NF=72   # launches per forward (scripts/one_forward.py prints it; unchanged since r1c)
timeout 120 ncu --metrics $MET --clock-control none -k "$MINE" -s $((2 * NF)) -c $NF --csv \
    --log-file gpurun_out/${TAG}_fwd_launches.csv python scripts/one_forward.py > gpurun_out/${TAG}_ncu1.log 2>&1
# steps 1 and 2 launch 62 kernels of this set each (53 bn_apply + 3 head fwd + 3 head bwd + 6 gathers = 65)
timeout 120 ncu --metrics $MET --clock-control none -k "$CHANGED" -s 130 -c 65 --csv \
    --log-file gpurun_out/${TAG}_train_changed_launches.csv python scripts/one_train_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log
wc -l gpurun_out/${TAG}_fwd_launches.csv gpurun_out/${TAG}_train_changed_launches.csv
