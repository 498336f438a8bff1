#!/bin/bash
# ON THE GPU BOX (gpurun): ncu evidence for the round.
#  1. launch list (time + DRAM bytes + tensor pipe) of every librotmv kernel of one warm inference forward
#  2. the same for one warm training step
#  3. --set full for the first 30 librotmv kernels of the warm forward (stem, max-pool, layer1, layer2, layer3 head)
set -u
TAG=${1:-r1c}
MINE='regex:igemm_kernel|wgrad_kernel|stem_|maxpool|avgpool|rotate_gather|head_loss|bn_|relu_bwd|colsum|permute_cast|adam_|dilate|simt_'
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
python scripts/one_forward.py > gpurun_out/${TAG}_fwd_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_fwd_plain.log; exit 1; }
python scripts/one_train_step.py > gpurun_out/${TAG}_train_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_train_plain.log; exit 1; }
NF=$(grep -m1 -o '[0-9]* launches' gpurun_out/${TAG}_fwd_plain.log | cut -d' ' -f1)
NT=$(grep -m1 -o '[0-9]* launches' gpurun_out/${TAG}_train_plain.log | cut -d' ' -f1)
echo "launches per forward: $NF, per training step: $NT"
ncu --metrics $MET --clock-control none -k "$MINE" -s $((2 * NF)) -c $NF --csv \
    --log-file gpurun_out/${TAG}_fwd_launches.csv python scripts/one_forward.py > gpurun_out/${TAG}_ncu1.log 2>&1
# the first training step launches one extra kernel per weight tensor (generic re-layout) -> skip by count of step 2
N1=$(sed -n 1p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
N2=$(sed -n 2p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
N3=$(sed -n 3p gpurun_out/${TAG}_train_plain.log | grep -o '[0-9]* launches' | cut -d' ' -f1)
ncu --metrics $MET --clock-control none -k "$MINE" -s $((N1 + N2)) -c $N3 --csv \
    --log-file gpurun_out/${TAG}_train_launches.csv python scripts/one_train_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k "$MINE" -s $((2 * NF)) -c 30 \
    -o /tmp/${TAG}_full python scripts/one_forward.py > gpurun_out/${TAG}_ncu3.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_full.ncu-rep gpurun_out | tail -8
tail -2 gpurun_out/${TAG}_fwd_plain.log gpurun_out/${TAG}_train_plain.log
