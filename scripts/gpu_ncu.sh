#!/bin/bash
# usage: gpu_ncu.sh <tag> <model> [n] [kernel-regex]
mkdir -p gpurun_out
TAG=$1; MODEL=${2:-tcp}; N=${3:-1048576}; K=${4:-ensemble_kernel}
python scripts/profile_run.py $MODEL $N > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_$TAG -f python scripts/profile_run.py $MODEL $N > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_$TAG.log
