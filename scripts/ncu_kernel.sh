#!/bin/bash
# ncu --set full capture of selected kernels of the train step.  Usage: bash scripts/ncu_kernel.sh tag kernel-regex [count] [skip]
TAG=$1; RE=$2; CNT=${3:-4}; SKIP=${4:-40}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} \
    python scripts/profile_fwd.py tf32 3 > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_${TAG}.log
