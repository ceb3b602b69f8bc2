#!/bin/bash
# A/B of kernel build variants (scripts/build_variant.py) on the C2 kernel at 4 M trajectories, policy 1
mkdir -p gpurun_out
V=piecewisedeterministicmarkovprocesses.jl_b200/variants
( echo "main"; PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-160
for v in "$@"; do echo "$v"; PDMP_B200_LIB=$PWD/$V/libpdmp_b200_$v.so PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-160; done ) | tee gpurun_out/variants.log
