#!/bin/bash
# ncu --set full capture of selected GEMM launches of one generator pass.  Usage: ncu_full.sh TAG SKIP COUNT
TAG=${1:-x}; SKIP=${2:-54}; COUNT=${3:-2}
mkdir -p gpurun_out
python scripts/ncu_target.py --which g > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sm100_kernel \
    -s $SKIP -c $COUNT -f -o gpurun_out/prof_$TAG python scripts/ncu_target.py --which g > gpurun_out/ncu_run_$TAG.log 2>&1
echo "ncu_exit=$?"; tail -3 gpurun_out/ncu_run_$TAG.log
