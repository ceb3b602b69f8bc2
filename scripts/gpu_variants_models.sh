#!/bin/bash
# A/B of build variants on the C3 (morris_lecar) and C5 (tcp2d) profile workloads
mkdir -p gpurun_out
V=piecewisedeterministicmarkovprocesses.jl_b200/variants
run() { python scripts/profile_run.py $1 $2 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-150; }
( echo main; run morris_lecar 1048576; run tcp2d 4194304
for v in "$@"; do echo $v; export PDMP_B200_LIB=$PWD/$V/libpdmp_b200_$v.so; run morris_lecar 1048576; run tcp2d 4194304; unset PDMP_B200_LIB; done ) | tee gpurun_out/variants_models.log
