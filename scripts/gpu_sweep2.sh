#!/bin/bash
mkdir -p gpurun_out
for cap in 2 3 4; do for jt in 12 16 20 24; do
  PDMP_JUMP_CAP=$cap PDMP_JUMP_THRESH=$jt PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-120 | sed -e "s/^/cap $cap /"
done; done | tee gpurun_out/sweep2_r2.log
for cap in 2 3; do for jt in 4 6 8 12; do
  PDMP_JUMP_CAP=$cap PDMP_JUMP_THRESH=$jt timeout 300 python scripts/profile_run.py morris_lecar 1048576 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-120 | sed -e "s/^/cap $cap /"
done; done | tee -a gpurun_out/sweep2_r2.log
