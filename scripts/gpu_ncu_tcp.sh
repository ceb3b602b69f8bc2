#!/bin/bash
# one --set full capture of the C2 ensemble kernel at 4 M trajectories (policy 1); usage: gpu_ncu_tcp.sh TAG [model] [n]
TAG=${1:-x}; MODEL=${2:-tcp}; N=${3:-4194304}
mkdir -p gpurun_out
PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py $MODEL $N > gpurun_out/${TAG}_plain.log 2>&1 && \
PDMP_TSTOP_POLICY=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof -f python scripts/profile_run.py $MODEL $N > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/${TAG}_plain.log | cut -c1-400
