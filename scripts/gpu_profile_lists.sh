#!/bin/bash
# ON THE GPU BOX (gpurun): ncu launch lists only (time, DRAM bytes, tensor-pipe activity) of every
# librotmv kernel of one warm inference forward and one warm training step.
set -u
TAG=${1:-r1d}
MINE='regex:igemm_|wgrad_|stem_|maxpool|avgpool|rotate_gather|head_loss|bn_|relu_bwd|colsum|permute_cast|adam_|dilate|simt_'
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
python scripts/one_forward.py > gpurun_out/${TAG}_fwd_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_fwd_plain.log; exit 1; }
python scripts/one_train_step.py > gpurun_out/${TAG}_train_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_train_plain.log; exit 1; }
NF=$(grep -m1 -o '[0-9]* launches' gpurun_out/${TAG}_fwd_plain.log | cut -d' ' -f1)
N1=$(sed -n 1p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
N2=$(sed -n 2p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
N3=$(sed -n 3p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
echo "launches: forward $NF, training steps $N1 $N2 $N3"
ncu --metrics $MET --clock-control none -k "$MINE" -s $((2 * NF)) -c $NF --csv \
    --log-file gpurun_out/${TAG}_fwd_launches.csv python scripts/one_forward.py > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --metrics $MET --clock-control none -k "$MINE" -s $((N1 + N2)) -c $N3 --csv \
    --log-file gpurun_out/${TAG}_train_launches.csv python scripts/one_train_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
cat gpurun_out/${TAG}_fwd_plain.log gpurun_out/${TAG}_train_plain.log
