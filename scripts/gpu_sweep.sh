#!/bin/bash
# scheduler-threshold sweep of the C2 kernel at 4 M trajectories (policy 1)
mkdir -p gpurun_out
for jt in 12 16 20 24 28; do for rt in 2 4 8; do
  PDMP_JUMP_THRESH=$jt PDMP_REFILL_THRESH=$rt PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200
done; done | tee gpurun_out/sweep_r2.log
