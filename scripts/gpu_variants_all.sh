#!/bin/bash
# A/B of build variants on the TCP (4 M, policy 1), Morris-Lecar and tcp2d profile workloads; usage: gpu_variants_all.sh V1 V2 ...
mkdir -p gpurun_out
V=piecewisedeterministicmarkovprocesses.jl_b200/variants
run() { PDMP_TSTOP_POLICY=$3 python scripts/profile_run.py $1 $2 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-110; }
all() { run tcp 4194304 1; run morris_lecar 1048576 0; run tcp2d 4194304 0; }
( echo main; all
for v in "$@"; do echo $v; export PDMP_B200_LIB=$PWD/$V/libpdmp_b200_$v.so; all; unset PDMP_B200_LIB; done ) | tee gpurun_out/variants_all.log
