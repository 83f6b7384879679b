#!/bin/bash
# ncu launch list of ONE embed+detect+locate step (64 x 1 s): per-launch device time and DRAM bytes.
TAG=${1:-x}
mkdir -p gpurun_out
python scripts/ncu_target.py > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python scripts/ncu_target.py > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu_exit=$?"; tail -2 gpurun_out/ncu_launches_$TAG.log; wc -l gpurun_out/launches_$TAG.csv
