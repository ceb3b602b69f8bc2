#!/bin/bash
# scheduler thresholds re-swept after the rcp / division changes (Morris-Lecar 2^20, tcp2d 2^22)
mkdir -p gpurun_out
for jt in 4 6 8 10 12 16; do
  PDMP_JUMP_THRESH=$jt timeout 300 python scripts/profile_run.py morris_lecar 1048576 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-100 | sed -e "s/^/JT $jt /"
done | tee gpurun_out/sweep3_r2.log
for rt in 4 8 16; do for jt in 8 12 16 20 24; do
  PDMP_REFILL_THRESH=$rt PDMP_JUMP_THRESH=$jt timeout 300 python scripts/profile_run.py tcp2d 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-100 | sed -e "s/^/RT $rt JT $jt /"
done; done | tee -a gpurun_out/sweep3_r2.log
